#!/bin/bash
python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r02f_gputests.log
tail -3 gpurun_out/r02f_gputests.log
B="python bench.py --steps 20 --warmup 3 --skip-cpu --skip-e2e --skip-configs --skip-sustained"
for th in 2048 4096 8192 16384 32768 65536; do
  for bl in 0 1; do
    RODEO_DALTON_BLOCK_LANES=$bl $B --thetas $th 2>gpurun_out/r02f_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('block_lanes=$bl', d['config']['thetas_per_gpu'], d['ms_per_step'], d['roofline']['kernel_ms'])"
  done
done
python bench.py --steps 20 --warmup 3 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; tail -3 gpurun_out/r02f_bench.err
