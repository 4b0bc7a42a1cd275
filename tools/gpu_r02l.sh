#!/bin/bash
python -m pytest tests -q -m gpu 2>&1 | tail -6
C="python tools/bench_configs.py --only C4 --reps 3"
$C > gpurun_out/plain_r02_C4.log 2>&1 && grep '^{' gpurun_out/plain_r02_C4.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'][:40],'ms',round(d['ms'],3),'frac',round(d['roofline_frac'],3))" &&
ncu --set full --clock-control none --import-source on -k regex:fenrir_ws_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02_C4 $C > gpurun_out/ncu_full_r02_C4.log 2>&1
{ ncu -i gpurun_out/prof_r02_C4.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_key_metrics.py; echo; ncu -i gpurun_out/prof_r02_C4.ncu-rep --page source --csv --print-source sass 2>/dev/null | python tools/ncu_source_summary.py $((16384*2000/32)); } > gpurun_out/summary_r02_C4.txt 2>&1
rm -f gpurun_out/prof_r02_C4.ncu-rep
head -12 gpurun_out/summary_r02_C4.txt | cut -c1-140
RODEO_FENRIR_WS=1 python - <<'PY'
# large-batch check of the warp-specialised kernel against the one-warp kernel (FN, 65,536 thetas)
import os, sys, time, numpy as np, torch
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import problems as P, rodeo_b200 as rb
pr = P.fitz_problem(65536, seed=0); ob = P.fitz_obs(pr, None)
X0, th = torch.as_tensor(pr["X0"]).cuda(), torch.as_tensor(pr["theta"]).cuda()
def run():
    return rb.inference.fenrir(None, rb.models.fitzhugh_nagumo, pr["W"], X0, 0.0, 40.0, 800, rb.interrogate.interrogate_kramer, prior_pars=(pr["Q"], pr["R"]), theta=th, **ob)
out = {}
for ws in ("1", "0"):
    os.environ["RODEO_FENRIR_WS"] = ws
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = run(); e1.record(); torch.cuda.synchronize()
    out[ws] = r; print("FN fenrir 65,536 thetas WS=%s: %.3f ms" % (ws, e0.elapsed_time(e1)))
print("bitwise equal:", bool(torch.equal(out["0"], out["1"])))
PY
