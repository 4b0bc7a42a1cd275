#!/bin/bash
export RODEO_DALTON_BLOCK_LANES=1
CMD="python bench.py --thetas 4096 --steps 2 --warmup 3 --skip-cpu --skip-e2e"
$CMD > gpurun_out/r02e_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dalton_bl_kernel -s 3 -c 1 -f -o gpurun_out/prof_dalton_bl_small $CMD > gpurun_out/r02e_ncu.log 2>&1
tail -2 gpurun_out/r02e_ncu.log
