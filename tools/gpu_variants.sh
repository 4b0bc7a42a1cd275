python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "float32" 2>&1 | tail -12
python - <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, "tests")
import problems as P, rodeo_b200 as rb
pr = P.fitz_problem(65536, seed=0); ob = P.fitz_obs(pr, None)
kr = rb.interrogate.interrogate_kramer
for dt in (np.float64, np.float32):
    X0 = torch.as_tensor(pr["X0"].astype(dt)).cuda(); th = torch.as_tensor(pr["theta"].astype(dt)).cuda()
    f = lambda: rb.inference.dalton(None, rb.models.fitzhugh_nagumo, pr["W"], X0, 0.0, 40.0, 800, kr, prior_pars=(pr["Q"], pr["R"]), theta=th, **ob)
    for _ in range(3): f()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); [f() for _ in range(10)]; e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(dt.__name__, "dalton ms", round(ms, 3), "G theta-steps/s", round(65536 * 800 / ms / 1e6, 1))
PY
