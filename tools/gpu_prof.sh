#!/bin/bash
# ncu --set full of one kernel of tools/bench_configs.py, summarised on the box (the .ncu-rep stays there):
#   bash tools/gpu_prof.sh <config> <kernel-regex> <tag> <units>     e.g.  C1 solve_mv_bl_kernel c1 $((65536*800/32))
# writes gpurun_out/summary_<tag>.txt (key metrics + SASS opcode mix), source_<tag>.csv (per-SASS-line), lines_<tag>.csv
# (per-CUDA-line)
CMD="python tools/bench_configs.py --only $1 --reps 1"
$CMD > gpurun_out/plain_$3.log 2>&1 || { tail -5 gpurun_out/plain_$3.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o gpurun_out/prof_$3 $CMD > gpurun_out/ncu_full_$3.log 2>&1
{ ncu -i gpurun_out/prof_$3.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_key_metrics.py; echo
  ncu -i gpurun_out/prof_$3.ncu-rep --page source --csv --print-source sass 2>/dev/null | python tools/ncu_source_summary.py $4; } > gpurun_out/summary_$3.txt 2>&1
ncu -i gpurun_out/prof_$3.ncu-rep --page source --csv --print-source sass > gpurun_out/source_$3.csv 2>/dev/null
ncu -i gpurun_out/prof_$3.ncu-rep --page source --csv --print-source cuda > gpurun_out/lines_$3.csv 2>/dev/null
rm -f gpurun_out/prof_$3.ncu-rep
head -24 gpurun_out/summary_$3.txt | cut -c1-150
