#!/bin/bash
python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r02d_gputests.log
tail -3 gpurun_out/r02d_gputests.log
B="python bench.py --steps 20 --warmup 3 --skip-cpu --skip-e2e"
for th in 2048 4096 8192 16384 32768 65536; do
  for bl in 0 1; do
    RODEO_DALTON_BLOCK_LANES=$bl $B --thetas $th 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('block_lanes=$bl', d['config']['thetas_per_gpu'], d['ms_per_step'])"
  done
done
python tools/bench_configs.py > gpurun_out/r02d_configs.log 2>&1; grep '^{' gpurun_out/r02d_configs.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'][:52],'ms',round(d['ms'],3),'frac',round(d['roofline_frac'],3))"
bash tools/gpu_profile_cfg.sh C4 fenrir_kernel r02d_c4
