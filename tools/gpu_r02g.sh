#!/bin/bash
python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/r02g_gputests.log
tail -8 gpurun_out/r02g_gputests.log
for ws in 0 1; do
  RODEO_FENRIR_WS=$ws python tools/bench_configs.py --only C4 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('WS=$ws', d['config'][:40],'ms',round(d['ms'],3),'frac',round(d['roofline_frac'],3))"
done
python tools/bench_configs.py --only C5,C5x 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'][:60],'ms',round(d['ms'],3),'frac',round(d['roofline_frac'],3))"
python bench.py --steps 20 --warmup 3 --skip-configs > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; tail -3 gpurun_out/r02g_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02g_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'e2e_py',d['e2e_python']['value'])"
