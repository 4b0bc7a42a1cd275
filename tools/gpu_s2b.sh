#!/bin/bash
# ncu --set full of the schedule kernels on C5 (log-lik only) and C3 (draws written)
prof() {  # config kernel-regex tag units
  CMD="python tools/bench_configs.py --only $1 --reps 1"
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o gpurun_out/prof_$3 $CMD > gpurun_out/ncu_full_$3.log 2>&1
  { ncu -i gpurun_out/prof_$3.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_key_metrics.py; echo
    ncu -i gpurun_out/prof_$3.ncu-rep --page source --csv --print-source sass 2>/dev/null | python tools/ncu_source_summary.py $4; } > gpurun_out/summary_$3.txt 2>&1
  ncu -i gpurun_out/prof_$3.ncu-rep --page source --csv --print-source sass > gpurun_out/source_$3.csv 2>/dev/null
  rm -f gpurun_out/prof_$3.ncu-rep
}
prof C5 solve_sim_sched_kernel s2_C5 $((32768*800/32))
prof C3 solve_sim_sched_kernel s2_C3 $((65536*4000/32))
head -30 gpurun_out/summary_s2_C5.txt | cut -c1-150
