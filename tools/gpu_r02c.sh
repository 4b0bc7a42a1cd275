#!/bin/bash
# ncu capture of the small-batch (latency-bound) dalton launch
export RODEO_DALTON_BLOCK_LANES=0
CMD="python bench.py --thetas 4096 --steps 2 --warmup 3 --skip-cpu --skip-e2e"
$CMD > gpurun_out/r02c_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dalton_kernel -s 3 -c 1 -f -o gpurun_out/prof_dalton_small $CMD > gpurun_out/r02c_ncu.log 2>&1
tail -3 gpurun_out/r02c_ncu.log
