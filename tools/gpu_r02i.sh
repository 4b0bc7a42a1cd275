#!/bin/bash
python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/r02i_gputests.log
tail -8 gpurun_out/r02i_gputests.log
for ws in 0 1; do
  RODEO_FENRIR_WS=$ws python tools/bench_configs.py --only C4 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('WS=$ws', d['config'][:40],'ms',round(d['ms'],3),'frac',round(d['roofline_frac'],3))"
done
