python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "injected or deterministic or pseudo or ragged or adaptive" 2>&1 | tail -3
python tools/bench_configs.py --only C5 2>&1 | cut -c1-200
