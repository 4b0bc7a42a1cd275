#!/bin/bash
# N GPUs (usage: gpurun --gpus N -- bash tools/gpu_scaling.sh N): hardware multi-rank test, then the benchmark as the driver launches it
N=${1:-2}
python -m pytest tests/test_multi_gpu.py -q -m gpu 2>&1 | tail -5
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then
      python bench.py --gpus 1 --steps 30 --warmup 3 --skip-configs --skip-sustained > gpurun_out/scale_bench_n1.json 2> gpurun_out/scale_bench_n1.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 30 --warmup 3 --skip-configs --skip-sustained > gpurun_out/scale_bench_n$n.json 2> gpurun_out/scale_bench_n$n.err
    fi
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/scale_bench_n$n.json').read().strip().splitlines()[-1])
    s=d.get('strong') or {}
    print('N=$n value %.4g ms %.4f e2e %.4g strong: ms %s speedup %s bitwise %s gather_ok %s' % (d['value'], d['ms_per_step'], d['e2e']['value'] or 0, s.get('ms_per_step'), s.get('speedup_vs_one_gpu'), s.get('sharded_equals_unsharded_bitwise'), d.get('gather_matches_rank_outputs')))
except Exception as e:
    print('N=$n failed', e); print(open('gpurun_out/scale_bench_n$n.err').read()[-1500:])
PY
  fi
done
