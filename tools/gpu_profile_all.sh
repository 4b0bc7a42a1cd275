# End-of-round evidence: bench launch list + ncu --set full of every config's kernel (each after its plain run exited 0).
# The reports are summarised ON THE BOX (gpurun_out/ may bring back at most 64 MiB); only the headline kernel's
# .ncu-rep is kept.       usage (on the GPU box): bash tools/gpu_profile_all.sh <tag>
TAG=${1:-r01}
summarise() {   # <report> <units for per-unit instruction counts> <out>
  { ncu -i $1 --page raw --csv 2>/dev/null | python tools/ncu_key_metrics.py
    echo
    ncu -i $1 --page source --csv --print-source sass 2>/dev/null | python tools/ncu_source_summary.py $2
  } > $3 2>&1
}
bash tools/gpu_profile.sh $TAG > gpurun_out/profile_all_$TAG.log 2>&1
summarise gpurun_out/prof_dalton_$TAG.ncu-rep $((2*2*65536*800/32)) gpurun_out/summary_${TAG}_dalton.txt
for spec in "C1 solve_mv_bl $((65536*800/32))" "C5 solve_sim_bl $((2*32768*800/32))" "C4 fenrir_kernel $((16384*2000/32))" "C3 solve_sim_kernel $((65536*4000/32))"; do
  set -- $spec
  bash tools/gpu_profile_cfg.sh $1 $2 ${TAG}_$1 >> gpurun_out/profile_all_$TAG.log 2>&1
  summarise gpurun_out/prof_${TAG}_$1.ncu-rep $3 gpurun_out/summary_${TAG}_$1.txt
  rm -f gpurun_out/prof_${TAG}_$1.ncu-rep
done
python tools/bench_configs.py > gpurun_out/bench_configs_$TAG.log 2>&1
python tools/bench_configs.py --only C1f32,C2f32 >> gpurun_out/bench_configs_$TAG.log 2>&1
tail -3 gpurun_out/profile_all_$TAG.log; du -sh gpurun_out
