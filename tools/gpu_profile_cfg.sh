# ncu --set full of one kernel of tools/bench_configs.py; usage: bash tools/gpu_profile_cfg.sh <config> <kernel-regex> <tag>
CFG=$1; KRN=$2; TAG=$3
CMD="python tools/bench_configs.py --only $CFG --reps 1"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KRN -s 2 -c 1 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
