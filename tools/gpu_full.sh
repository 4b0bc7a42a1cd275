python -m pytest tests -q -m gpu 2>&1 | tail -5
python tools/bench_configs.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:52],'ms',round(d['ms'],3),'G/s',round(d['theta_steps_per_s']/1e9,2),'frac',round(d['roofline_frac'],3), d['bound'])
    else: print(l.rstrip())
"
