# ncu evidence for the headline kernel; usage: bash tools/gpu_profile.sh <tag>
TAG=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dalton_kernel -s 3 -c 2 -o gpurun_out/prof_dalton_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out
