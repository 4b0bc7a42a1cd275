#!/bin/bash
# Round-2 evidence: bench launch list + ncu --set full of the headline kernel and of every config's kernel (each ncu
# run directly after the same command exited 0 without ncu).  Summaries are made ON THE BOX; only the headline kernel's
# report is kept.        usage (on the GPU box): bash tools/gpu_profile_r02.sh
TAG=r02
summarise() {   # <report> <units for per-unit instruction counts> <out>
  { ncu -i $1 --page raw --csv 2>/dev/null | python tools/ncu_key_metrics.py
    echo
    ncu -i $1 --page source --csv --print-source sass 2>/dev/null | python tools/ncu_source_summary.py $2
  } > $3 2>&1
}
CMD="python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-configs --skip-sustained"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dalton_kernel -s 3 -c 1 -f -o gpurun_out/prof_dalton_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
summarise gpurun_out/prof_dalton_$TAG.ncu-rep $((65536*800/32)) gpurun_out/summary_${TAG}_dalton.txt
for spec in "C1 solve_mv_bl $((65536*800/32))" "C5 solve_sim_sched_kernel $((32768*800/32))" "C4 fenrir_ws_kernel $((16384*2000/32))" "C3 solve_sim_sched_kernel $((65536*4000/32))"; do
  set -- $spec
  C="python tools/bench_configs.py --only $1 --reps 1"
  $C > gpurun_out/plain_${TAG}_$1.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_$1 $C > gpurun_out/ncu_full_${TAG}_$1.log 2>&1
  summarise gpurun_out/prof_${TAG}_$1.ncu-rep $3 gpurun_out/summary_${TAG}_$1.txt
  rm -f gpurun_out/prof_${TAG}_$1.ncu-rep
done
python tools/bench_configs.py > gpurun_out/bench_configs_$TAG.log 2>&1
python tools/bench_configs.py --only C1f32,C2f32,C5x >> gpurun_out/bench_configs_$TAG.log 2>&1
python bench.py --steps 30 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
grep -h "^{" gpurun_out/bench_configs_$TAG.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'][:56],'ms',round(d['ms'],3),'frac',round(d['roofline_frac'],3))"
head -12 gpurun_out/summary_${TAG}_dalton.txt | cut -c1-160
du -sh gpurun_out
