"""The NumPy stand-in for jax (oracle/jaxshim) that produced tests/golden/reference_vectors.npz: the JAX semantics it
emulates on purpose are checked here, independently of the reference.  (That the reference's own
test-suite passes over it is checked in the build container by tests/golden/run_reference_tests_over_shim.py.)"""
import importlib
import os
import sys

import numpy as np
import pytest

SHIM = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "jaxshim")


@pytest.fixture(scope="module")
def jax():
    assert "jax" not in sys.modules or "numpy-shim" in getattr(sys.modules["jax"], "__version__", ""), \
        "a real jax is already imported"
    sys.path.insert(0, SHIM)
    try:
        mod = importlib.import_module("jax")
        importlib.import_module("jax.numpy")
        yield mod
    finally:
        sys.path.remove(SHIM)
        for name in [m for m in sys.modules if m == "jax" or m.startswith("jax.")]:
            del sys.modules[name]


def test_indexing_never_raises_wraps_then_clamps(jax):
    jnp = jax.numpy
    a = jnp.arange(5) * 10
    assert int(a[7]) == 40 and int(a[-1]) == 40 and int(a[-9]) == 0        # clamp high, wrap, clamp low
    m = jnp.arange(12).reshape(3, 4)
    assert int(m[5, 1]) == 9 and list(m[-1]) == [8, 9, 10, 11]
    assert [int(v) for v in a] == [0, 10, 20, 30, 40]                       # iteration still terminates
    b = a.at[1].set(-3)
    assert int(b[1]) == -3 and int(a[1]) == 10                              # functional update


def test_scan_vmap_cond_follow_jax(jax):
    jnp, lax = jax.numpy, jax.lax
    xs = {"t": jnp.arange(4), "v": jnp.arange(8.0).reshape(4, 2)}
    carry, ys = lax.scan(lambda c, x: (c + x["t"], (c, x["v"].sum())), 0, xs)
    assert int(carry) == 6 and [int(v) for v in ys[0]] == [0, 0, 1, 3] and ys[1].shape == (4,)
    carry, ys = lax.scan(lambda c, x: (c + x, c), 0, jnp.arange(4), reverse=True)
    assert int(carry) == 6 and [int(v) for v in ys] == [6, 5, 3, 0]        # run from the end, stacked at x's own index
    f = lambda a, b: (a @ b, a.sum())
    A, Bm = np.random.default_rng(0).standard_normal((3, 2, 2)), np.random.default_rng(1).standard_normal((3, 2))
    out = jax.vmap(f)(a=jnp.array(A), b=jnp.array(Bm))
    assert np.allclose(out[0], np.einsum("nij,nj->ni", A, Bm)) and out[1].shape == (3,)
    out = jax.vmap(lambda k, shape: jnp.zeros(shape) + k, in_axes=(0, None))(jnp.arange(3), (2,))
    assert out.shape == (3, 2)
    assert lax.cond(jnp.array(3) == 3, lambda: 1, lambda: 0) == 1


def test_jacfwd_is_exact_to_rounding_and_linalg_matches_numpy(jax):
    jnp = jax.numpy
    th = np.array([0.2, 0.2, 3.0])

    def fitz(X, t, theta):
        a, b, c = theta
        V, R = X[:, 0]
        return jnp.array([[c * (V - V * V * V / 3 + R)], [-1 / c * (V - a + b * R)]])
    X = np.array([[-0.7, 0.3, 0.1], [0.9, -0.2, 0.0]])
    J = np.asarray(jax.jacfwd(fitz)(jnp.array(X), 0.0, th))
    assert J.shape == (2, 1, 2, 3)
    want = np.zeros((2, 1, 2, 3))
    want[0, 0, 0, 0] = 3.0 * (1 - 0.49); want[0, 0, 1, 0] = 3.0
    want[1, 0, 0, 0] = -1 / 3.0; want[1, 0, 1, 0] = -0.2 / 3.0
    assert np.max(np.abs(J - want)) < 1e-15
    S = np.array([[2.0, 0.3], [0.3, 1.0]])
    assert np.allclose(jnp.linalg.cholesky(jnp.array(S + np.array([[0, 1e-3], [-1e-3, 0]]))), np.linalg.cholesky(S))
    assert np.isnan(np.asarray(jnp.linalg.cholesky(jnp.array([[1.0, 2.0], [2.0, 1.0]])))).all()     # NaN, not an exception


def test_random_is_keyed_and_logged(jax):
    r = jax.random
    k = r.PRNGKey(3)
    ks = r.split(k, 4)
    assert ks.shape == (4, 2) and len({tuple(int(v) for v in row) for row in np.asarray(ks)}) == 4
    del r.DRAW_LOG[:]
    z1 = np.asarray(r.normal(ks[0], (3,)))
    z2 = np.asarray(r.normal(ks[0], (3,)))
    assert np.array_equal(z1, z2) and len(r.DRAW_LOG) == 2 and np.array_equal(r.DRAW_LOG[0][2], z1)
    cov = np.array([[2.0, 0.5], [0.5, 1.0]])
    x = np.asarray(r.multivariate_normal(ks[1], np.zeros(2), cov, method="svd"))
    u, s, _ = np.linalg.svd(cov)
    assert np.allclose(x, (u * np.sqrt(s)) @ r.DRAW_LOG[-1][2])
