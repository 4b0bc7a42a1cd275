# usage: bash tools/gpu_variants_bench.sh <lib> [<lib> ...]   -- headline bench (device-resident leg only) on tuning builds
for L in "$@"; do echo "== $L"; RODEO_B200_LIB=$L python bench.py --steps 30 --warmup 3 --skip-cpu --skip-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms',round(d['ms_per_step'],4),'G/s',round(d['value']/1e9,2))
"; done
