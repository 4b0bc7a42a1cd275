"""CPU: the C-ABI library loads without a GPU and exports every symbol include/rodeo_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rodeo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rodeo_b200_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ("rodeo_b200_solve_mv_f64", "rodeo_b200_solve_sim_f64", "rodeo_b200_dalton_f64",
              "rodeo_b200_fenrir_f64", "rodeo_b200_basic_gather_f64", "rodeo_b200_ode_init_pad_f64",
              "rodeo_b200_workspace_bytes", "rodeo_b200_last_error", "rodeo_b200_dalton_f64_host"):
        assert s in syms


def test_library_loads_and_exports_every_declared_symbol():
    from rodeo_b200 import _lib
    lib = _lib.load()
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/rodeo_b200.h but not exported"
    assert lib.rodeo_b200_abi_version() == 1
    # every declared symbol also has a ctypes signature in the binding
    assert set(declared_symbols()) <= set(_lib.SIGNATURES) | {"rodeo_b200_abi_version"}


def test_struct_layout_matches_header():
    from rodeo_b200 import _lib
    # 2 x int64 + 10 x int32 + 2 x uint32 + 2 x double, naturally aligned
    # + user_wcol, prior_batched, prior_var_scale
    assert ctypes.sizeof(_lib.RodeoProblem) == 16 + 40 + 8 + 16 + 8 + 8


def test_struct_size_matches_the_library_and_the_documented_binding():
    """sizeof(RodeoProblem) as compiled == the ctypes mirror in rodeo_b200/_lib.py == the struct INTEGRATION.md tells
    a binding author to declare (a short struct would hand the library a garbage prior_var_scale pointer)."""
    from rodeo_b200 import _lib
    lib = _lib.load()
    n = lib.rodeo_b200_problem_sizeof()
    assert n == ctypes.sizeof(_lib.RodeoProblem)
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"class RodeoProblem\(ctypes\.Structure\):.*?\n\n", doc, flags=re.S)
    assert m, "INTEGRATION.md no longer shows the ctypes struct"
    ns = {"ctypes": ctypes}
    exec(m.group(0), ns)
    assert ctypes.sizeof(ns["RodeoProblem"]) == n
    assert [f[0] for f in ns["RodeoProblem"]._fields_] == [f[0] for f in _lib.RodeoProblem._fields_]


def test_workspace_sizing_needs_no_gpu():
    from rodeo_b200 import _lib
    lib = _lib.load()
    c = _lib.RodeoProblem()
    c.B, c.n_steps, c.n_block, c.n_bstate = 65536, 800, 2, 3
    # solve_mv ((theta, block)-lane kernel): one checkpoint of n_block*(p + p(p+1)/2) = 18 doubles per K = 9 steps
    # (ceil(800/9) - 1 = 88 of them)
    n = lib.rodeo_b200_workspace_bytes(_lib.OP_SOLVE_MV, ctypes.byref(c), 8)
    assert n == 88 * 18 * 65536 * 8
    # solve_sim / fenrir: every filtered state filt[1..N-1]
    assert lib.rodeo_b200_workspace_bytes(_lib.OP_SOLVE_SIM, ctypes.byref(c), 8) == 799 * 18 * 65536 * 8
    assert lib.rodeo_b200_workspace_bytes(_lib.OP_FENRIR, ctypes.byref(c), 8) == 799 * 18 * 65536 * 8
    assert lib.rodeo_b200_workspace_bytes(_lib.OP_DALTON, ctypes.byref(c), 8) == 0


def test_compute_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    import rodeo_b200
    with pytest.raises(Exception) as e:
        rodeo_b200.solve_mv(None, rodeo_b200.models.fitzhugh_nagumo, np.zeros((2, 1, 3)), np.zeros((2, 3)), 0.0, 1.0,
                            4, rodeo_b200.interrogate.interrogate_kramer, prior_pars=(np.zeros((2, 3, 3)),) * 2,
                            theta=np.ones(3))
    assert "no CPU fallback" in str(e.value)
