#!/bin/bash
# round-2 experiment: new tests, dalton geometries at several batch sizes, register-cap variant
set -x
python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r02b_gputests.log
tail -3 gpurun_out/r02b_gputests.log
B="python bench.py --steps 20 --warmup 3 --skip-cpu --skip-e2e"
for th in 65536; do
  $B --thetas $th 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default lib', d['config']['thetas_per_gpu'], d['ms_per_step'])"
  RODEO_B200_LIB=build/variants/minb12.so $B --thetas $th 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('minb12', d['config']['thetas_per_gpu'], d['ms_per_step'])"
done
for th in 4096 8192 16384 24576 32768 65536; do
  for bl in 0 1; do
    RODEO_DALTON_BLOCK_LANES=$bl $B --thetas $th 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('block_lanes=$bl', d['config']['thetas_per_gpu'], d['ms_per_step'])"
  done
done
