python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -4
for b in 32768 65536 75776; do
  echo "== B=$b"
  python bench.py --steps 20 --warmup 3 --skip-cpu --thetas $b 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step',d['ms_per_step'],'value G',d['value']/1e9, 'e2e G', d['e2e']['value']/1e9, 'frac', d['roofline']['frac'])
    else: print(l.rstrip())
"
done
