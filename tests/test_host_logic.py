"""CPU: host-side mirror of the reference interface (no compute)."""
import functools

import numpy as np
import pytest

import rodeo_b200
from rodeo_b200 import _lib, interrogate, models, prior
from oracle import rodeo_oracle as orc


def test_public_names_mirror_the_reference():
    # reference src/rodeo/__init__.py:1-6 and src/rodeo/inference/__init__.py:1-3
    for name in ("solve_mv", "solve_sim", "interrogate", "prior", "inference", "utils"):
        assert hasattr(rodeo_b200, name)
    for name in ("basic", "fenrir", "dalton", "magi_logdens"):
        assert hasattr(rodeo_b200.inference, name)
    for name in ("interrogate_kramer", "interrogate_chkrebtii", "interrogate_schober", "interrogate_rodeo"):
        assert hasattr(rodeo_b200.interrogate, name)
    assert hasattr(rodeo_b200.prior, "ibm_init") and hasattr(rodeo_b200.utils, "first_order_pad")
    # reference src/rodeo/prior/__init__.py:1-2
    Q, R = rodeo_b200.prior.ibm_init(0.1, 3, [0.1, 0.2])
    Qd, Rd = rodeo_b200.prior.indep_init((Q, R))
    assert Qd.shape == (1, 6, 6) and np.array_equal(Qd[0, 3:, 3:], Q[1]) and not Qd[0, :3, 3:].any()
    assert np.array_equal(Rd[0, :3, :3], R[0])


def test_interrogation_objects_resolve_by_identity():
    assert interrogate.resolve(interrogate.interrogate_kramer) == _lib.INTERROGATE_KRAMER
    assert interrogate.resolve(interrogate.interrogate_schober) == _lib.INTERROGATE_SCHOBER
    assert interrogate.resolve(interrogate.interrogate_rodeo) == _lib.INTERROGATE_RODEO
    part = functools.partial(interrogate.interrogate_chkrebtii, kalman_type="standard")
    assert interrogate.resolve(part) == _lib.INTERROGATE_CHKREBTII
    with pytest.raises(TypeError):            # the reference fails the same way: kalman_type is never passed
        interrogate.resolve(interrogate.interrogate_chkrebtii)
    with pytest.raises(NotImplementedError):  # unknown callables: no CPU fallback
        interrogate.resolve(lambda *a, **k: None)


def test_models_resolve_and_reject_python_callables():
    assert models.resolve(models.fitzhugh_nagumo).model_id == 0
    assert models.resolve("lorenz63").n_block == 3
    with pytest.raises(NotImplementedError):
        models.resolve(lambda X, t, **p: X)


def test_ibm_init_equals_oracle_and_is_unit_upper():
    for dt, p, sig in ((0.05, 3, [0.1, 0.1]), (0.005, 4, [0.001]), (0.005, 3, [5e7] * 3)):
        Q, R = prior.ibm_init(dt, p, sig)
        Qo, Ro = orc.ibm_init(dt, p, np.array(sig))
        assert np.array_equal(Q, Qo) and np.array_equal(R, Ro)
        assert np.all(np.diagonal(Q, axis1=1, axis2=2) == 1.0) and not np.tril(Q, -1).any()


def test_kalman_type_errors_like_the_reference():
    from rodeo_b200 import _host
    assert _host.kalman_id("standard") == 0 and _host.kalman_id("square-root") == 1
    with pytest.raises(NotImplementedError):
        _host.kalman_id("bogus")                # reference src/rodeo/solve.py:236-241


def test_prior_spellings():
    from rodeo_b200 import _host
    Q, R = np.eye(3)[None], np.eye(3)[None]
    a = _host.prior_from((Q, R), None, None)
    b = _host.prior_from(None, Q, R)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    with pytest.raises(TypeError):
        _host.prior_from((Q, R), Q, None)
    with pytest.raises(TypeError):
        _host.prior_from(None, None, None)


def test_key_parsing():
    from rodeo_b200 import _host
    assert _host.parse_key(None) == (0, 0)
    assert _host.parse_key(np.array([7, 9], dtype=np.uint32)) == (7, 9)
    assert _host.parse_key(5) == (0, 5)
    with pytest.raises(ValueError):
        _host.parse_key([1, 2, 3])


def test_host_input_marshalling_needs_no_device():
    """Problem(host_inputs=True) keeps theta / ode_init / observation arrays as zero-copy CPU tensors for the *_host entry
    points: the struct, the broadcasting and the per-theta prior scale are the same as on the device path."""
    import torch
    from rodeo_b200 import _host
    import problems as P
    B = 5
    pr = P.fitz_problem(B, n_steps=20, t_max=1.0, seed=3)
    ob = P.fitz_obs(pr, None, n_obs=3)
    assert _host.on_host(pr["theta"], pr["X0"], None, [1.0, 2.0], torch.zeros(2))
    pb = _host.Problem(None, models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 1.0, 20, interrogate.interrogate_kramer,
                       (pr["Q"], pr["R"]), None, None, "standard", {"theta": pr["theta"]}, host_inputs=True)
    pb.set_obs(ob["obs_data"], ob["obs_times"], ob["obs_weight"], ob["obs_var"])
    assert pb.B == B and pb.c.B == B and pb.c.n_obs == 3 and pb.c.n_bobs == 1 and pb.c.prior_batched == 0
    assert not pb.theta.is_cuda and not pb.x0.is_cuda and not pb.obs_data.is_cuda and not pb.obs_ind.is_cuda
    assert pb.theta.data_ptr() == pr["theta"].ctypes.data          # zero-copy view of the caller's array
    assert np.array_equal(pb.obs_ind.numpy(), orc.obs_index(0.0, 1.0, 20, ob["obs_times"]))
    # un-batched theta with batched ode_init broadcasts; sigma part of theta becomes a host scale array
    Q, Rb = prior.ibm_init(1.0 / 20, 3, 0.1 * np.ones((B, 2)) * np.arange(1, B + 1)[:, None])
    pb = _host.Problem(None, models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 1.0, 20, interrogate.interrogate_kramer,
                       (Q, Rb), None, None, "standard", {"theta": pr["theta"][0]}, host_inputs=True)
    assert pb.B == B and tuple(pb.theta.shape) == (B, 3) and not pb.r_scale.is_cuda
    assert np.allclose(pb.r_scale.numpy()[:, 0], np.arange(1, B + 1) ** 2)


def test_host_path_without_gpu_fails_loudly():
    """dalton() on NumPy inputs goes through rodeo_b200_dalton_f64_host: without a GPU it must raise, not fall back"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import problems as P
    pr = P.fitz_problem(3, n_steps=10, t_max=0.5, seed=1)
    ob = P.fitz_obs(pr, None, n_obs=2)
    with pytest.raises(Exception) as e:
        rodeo_b200.inference.dalton(None, models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 0.5, 10,
                                    interrogate.interrogate_kramer, prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"], **ob)
    assert "CUDA" in str(e.value) or "cuda" in str(e.value)
