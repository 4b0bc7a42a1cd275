TAG=r01j
summarise() { { ncu -i $1 --page raw --csv 2>/dev/null | python tools/ncu_key_metrics.py; echo; ncu -i $1 --page source --csv --print-source sass 2>/dev/null | python tools/ncu_source_summary.py $2; } > $3 2>&1; }
bash tools/gpu_profile.sh $TAG > gpurun_out/profile_all_$TAG.log 2>&1
summarise gpurun_out/prof_dalton_$TAG.ncu-rep $((2*65536*800/32)) gpurun_out/summary_${TAG}_dalton.txt
python bench.py --steps 30 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python tools/bench_configs.py --only C2,C2f32 > gpurun_out/bench_configs_$TAG.log 2>&1
head -30 gpurun_out/summary_${TAG}_dalton.txt | cut -c1-150
