for l in 32 16 8; do echo "fenrir lanes $l"; RODEO_FENRIR_LANES=$l python tools/bench_configs.py --only C4 --reps 5 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ',d['config'][:30],'B',d['B'],'ms',round(d['ms'],3))
"; done
echo default; python tools/bench_configs.py --only C4 --reps 5 2>&1 | grep -o '"ms": [0-9.]*'
python -m pytest tests -q -m gpu -k "fenrir or ragged or golden" 2>&1 | tail -2
