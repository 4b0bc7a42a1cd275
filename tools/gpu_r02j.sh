#!/bin/bash
for v in np2 np3m4; do
  RODEO_B200_LIB=build/variants/$v.so python tools/bench_configs.py --only C4 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$v', d['config'][:40],'ms',round(d['ms'],3),'frac',round(d['roofline_frac'],3))"
done
