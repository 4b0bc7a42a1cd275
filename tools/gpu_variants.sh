python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "solve_sim or philox or nvrtc" 2>&1 | tail -5
python tools/bench_configs.py --only C5 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:52],'ms',round(d['ms'],3),'G/s',round(d['theta_steps_per_s']/1e9,2),'frac',round(d['roofline_frac'],3), d['bound'])
    else: print(l.rstrip())
"
