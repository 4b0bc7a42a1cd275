# usage: bash tools/gpu_variants.sh <configs> <lib> [<lib> ...]   -- tools/bench_configs.py on tuning builds
CFG=$1; shift
for L in "$@"; do echo "== $L"; RODEO_B200_LIB=$L python tools/bench_configs.py --only $CFG 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:40],'ms',round(d['ms'],3),'G/s',round(d['theta_steps_per_s']/1e9,2),'frac',round(d['roofline_frac'],3))
"; done
