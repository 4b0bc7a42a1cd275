python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "dalton or fenrir or cutoff" 2>&1 | tail -3
python bench.py --steps 30 --warmup 3 --skip-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step',d['ms_per_step'],'value G',d['value']/1e9, 'e2e G', d['e2e']['value']/1e9, 'frac', d['roofline']['frac'])
    else: print(l.rstrip())
"
