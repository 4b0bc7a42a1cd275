for sc in 0.25 0.0625; do for bl in 0 1; do echo "scale $sc bl $bl"; RODEO_SIM_BLOCK_LANES=$bl python tools/bench_configs.py --only C3,C5 --scale $sc --reps 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ',d['config'][:30],'B',d['B'],'ms',round(d['ms'],3))
"; done; done
