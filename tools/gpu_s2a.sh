#!/bin/bash
# schedule path: tests, then C3 / C5 / C5x with and without the schedule
python -m pytest tests/test_gpu_schedule.py -q -x 2>&1 | tail -15
python -m pytest tests/test_gpu_parity.py -q -x -k "solve_sim or lorenz or pseudo_marginal or ragged or float32" 2>&1 | tail -5
for S in 0 1; do
  echo "RODEO_SIM_SCHEDULE=$S"
  RODEO_SIM_SCHEDULE=$S python tools/bench_configs.py --only C3,C5,C5x --reps 5 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:52],'ms',round(d['ms'],3),'G/s',round(d['theta_steps_per_s']/1e9,2),'frac',round(d['roofline_frac'],3), d['bound'])
    else: print(l.rstrip())
"
done
