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
