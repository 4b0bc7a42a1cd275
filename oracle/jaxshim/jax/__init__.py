"""
A NumPy stand-in for the parts of ``jax`` that mlysy/rodeo's hot path uses  --  TEST INFRASTRUCTURE ONLY.

Why it exists: the reference is pure JAX and JAX is not installed in this image (no network), so the reference's
own source could not be executed to pin the oracle.  With this package first on ``sys.path`` the UNMODIFIED files
under ``/root/reference/src/rodeo`` import and run: ``jax.numpy`` is NumPy (float64, the same LAPACK routines jaxlib's
CPU backend calls: getrf/getrs, syevd, gesdd, potrf, geqrf), ``jax.vmap`` / ``jax.lax.scan`` / ``jax.lax.cond`` are
eager Python loops with JAX's pytree and indexing semantics, ``jax.jacfwd`` is a complex-step derivative (exact to
rounding for the analytic right-hand sides used), and ``jax.random`` is a keyed counter-based generator that LOGS every
standard-normal array it hands out, so the same normals can be injected into the oracle and the CUDA kernels.
``tests/golden/make_reference_golden.py`` uses it to generate ``tests/golden/reference_vectors.npz``, and the
reference's own test-suite passes over it unmodified (``tests/golden/run_reference_tests_over_shim.py``: 28 of 28).

What it does NOT pin: XLA's own operation order / fusion (agreement with real JAX is expected at ~1e-13, not bitwise)
and JAX's threefry random streams (draws are compared on injected normals only).

Nothing under ``rodeo_b200/`` imports this package; it is never on ``sys.path`` outside the generator script.

JAX semantics emulated on purpose:
  * integer indexing never raises: negative indices wrap once, then out-of-range indices clamp (``obs_ind[i]`` past
    the last observation in dalton / before the first in fenrir);
  * ``x.at[idx].set(v)`` functional updates;
  * ``lax.scan`` over dict / tuple pytrees, forward and ``reverse=True``, stacking the per-step outputs;
  * ``vmap`` with positional / keyword arguments and ``in_axes`` entries of ``0`` or ``None``.
"""
import sys
import types

import numpy as _np
import scipy.linalg as _sla
import scipy.special as _ssp
import scipy.stats as _sst

__version__ = "0.0-numpy-shim"


# ---------------------------------------------------------------------------------------------------------------------
# array type: NumPy array + JAX's indexing rules + .at[]
# ---------------------------------------------------------------------------------------------------------------------
class Array(_np.ndarray):
    def __new__(cls, a):
        return _np.asarray(a).view(cls)

    def __array_finalize__(self, obj):
        pass

    def _norm_index(self, idx):
        def fix(k, n):
            k = int(k)
            if k < 0:
                k += n
            return min(max(k, 0), n - 1)

        def is_int(k):
            return isinstance(k, (int, _np.integer)) and not isinstance(k, (bool, _np.bool_)) or \
                (isinstance(k, _np.ndarray) and k.ndim == 0 and k.dtype.kind in "iu")

        if is_int(idx):
            return fix(idx, self.shape[0]) if self.ndim else idx
        if isinstance(idx, tuple) and not any(k is Ellipsis or k is None for k in idx):
            out, ax = [], 0
            for k in idx:
                if is_int(k) and ax < self.ndim:
                    out.append(fix(k, self.shape[ax]))
                else:
                    out.append(k)
                ax += 1
            return tuple(out)
        return idx

    def __getitem__(self, idx):
        return super().__getitem__(self._norm_index(idx))

    def __iter__(self):                       # ndarray iterates through __getitem__ until IndexError, which never comes
        if self.ndim == 0:
            raise TypeError("iteration over a 0-d array")
        return (self[i] for i in range(self.shape[0]))

    def __round__(self, ndigits=None):        # unittest's assertAlmostEqual rounds 0-d results
        return round(float(self), ndigits)

    @property
    def at(self):
        return _At(self)


class _At:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, idx):
        return _AtIdx(self.a, idx)


class _AtIdx:
    def __init__(self, a, idx):
        self.a, self.idx = a, idx

    def set(self, v):
        out = _np.array(self.a, copy=True)
        out[self.idx] = v
        return Array(out)

    def add(self, v):
        out = _np.array(self.a, copy=True)
        out[self.idx] += v
        return Array(out)


def _wrap(x):
    if isinstance(x, _np.ndarray):
        return x.view(Array)
    if isinstance(x, tuple):
        return tuple(_wrap(v) for v in x)
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    if isinstance(x, (_np.floating, _np.integer, _np.complexfloating, _np.bool_)):
        return _np.asarray(x).view(Array)
    return x


def _wrapping(f):
    def g(*a, **k):
        return _wrap(f(*a, **k))
    g.__name__ = getattr(f, "__name__", "f")
    return g


# ---------------------------------------------------------------------------------------------------------------------
# pytrees
# ---------------------------------------------------------------------------------------------------------------------
def _tree_map(f, *trees):
    t = trees[0]
    if isinstance(t, dict):
        return {k: _tree_map(f, *[x[k] for x in trees]) for k in t}
    if isinstance(t, (tuple, list)):
        return type(t)(_tree_map(f, *[x[i] for x in trees]) for i in range(len(t)))
    if t is None:
        return None
    return f(*trees)


def _tree_leaves(t):
    if isinstance(t, dict):
        return [l for k in t for l in _tree_leaves(t[k])]
    if isinstance(t, (tuple, list)):
        return [l for x in t for l in _tree_leaves(x)]
    if t is None:
        return []
    return [t]


def _stack(outs):
    return _tree_map(lambda *xs: Array(_np.stack([_np.asarray(x) for x in xs])), *outs)


# ---------------------------------------------------------------------------------------------------------------------
# transformations
# ---------------------------------------------------------------------------------------------------------------------
def vmap(fun, in_axes=0, out_axes=0):
    assert out_axes == 0

    def mapped(*args, **kwargs):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        assert len(axes) == len(args)
        n = None
        for a, ax in list(zip(args, axes)) + [(v, 0) for v in kwargs.values()]:
            if ax is None:
                continue
            assert ax == 0
            for leaf in _tree_leaves(a):
                m = _np.shape(leaf)[0]
                assert n is None or n == m, "vmap: inconsistent mapped axis sizes"
                n = m
        take = lambda a, ax, i: a if ax is None else _tree_map(lambda x: _wrap(_np.asarray(x)[i]), a)
        outs = [fun(*[take(a, ax, i) for a, ax in zip(args, axes)], **{k: take(v, 0, i) for k, v in kwargs.items()})
                for i in range(n)]
        return _stack(outs)
    return mapped


def jit(fun=None, **_kw):
    if fun is None:
        return lambda f: f
    return fun


def jacfwd(fun, argnums=0):
    """Jacobian by the complex-step derivative Im f(x + i h e_k) / h, h = 1e-30: no subtractive cancellation, so it is
    exact to rounding for analytic `fun` (polynomials, exp / log / sin / cos ...)."""
    h = 1e-30

    def jac(*args, **kwargs):
        x = _np.asarray(args[argnums], dtype=_np.float64)
        f0 = _np.asarray(fun(*args, **kwargs))
        out = _np.zeros(f0.shape + x.shape)
        for k in _np.ndindex(*x.shape):
            xc = x.astype(_np.complex128)
            xc[k] += 1j * h
            a = list(args)
            a[argnums] = Array(xc)
            out[(Ellipsis,) + k] = _np.imag(_np.asarray(fun(*a, **kwargs))) / h
        return Array(out)
    return jac


def jacrev(fun, argnums=0):
    return jacfwd(fun, argnums)


def grad(fun, argnums=0):
    """gradient of a scalar function: complex step where `fun` is analytic in NumPy's complex arithmetic, central
    differences otherwise (eigh / cholesky inside a log-density are not complex-analytic).  Only the reference's
    jit-vs-eager consistency tests use it."""
    cs = jacfwd(fun, argnums)

    def g(*args, **kwargs):
        try:
            out = cs(*args, **kwargs)
            if _np.all(_np.isfinite(_np.asarray(out))):
                return out
        except Exception:
            pass
        x = _np.asarray(args[argnums], dtype=_np.float64)
        out = _np.zeros(x.shape)
        for k in _np.ndindex(*x.shape):
            h = 1e-6 * max(1.0, abs(float(x[k])))
            vals = []
            for sgn in (1.0, -1.0):
                xp = x.copy()
                xp[k] += sgn * h
                a = list(args)
                a[argnums] = Array(xp)
                vals.append(float(_np.asarray(fun(*a, **kwargs))))
            out[k] = (vals[0] - vals[1]) / (2 * h)
        return Array(out)
    return g


class _Config:
    def update(self, *_a, **_k):
        pass


config = _Config()


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


# ---------------------------------------------------------------------------------------------------------------------
# jax.numpy
# ---------------------------------------------------------------------------------------------------------------------
def _linalg_solve(a, b):
    try:
        return _wrap(_np.linalg.solve(a, b))
    except _np.linalg.LinAlgError:          # exactly singular: LAPACK getrf info > 0; jax returns non-finite values
        return _wrap(_np.full(_np.broadcast_shapes(_np.shape(b)), _np.nan))


def _cholesky(a, symmetrize_input=True):
    a = _np.asarray(a)
    if symmetrize_input:                    # jnp.linalg.cholesky's default: factor (A + A^H) / 2
        a = 0.5 * (a + _np.swapaxes(a, -1, -2))
    try:
        return _wrap(_np.linalg.cholesky(a))
    except _np.linalg.LinAlgError:          # jax returns NaN for a non-PD input instead of raising
        return _wrap(_np.full(_np.shape(a), _np.nan))


_linalg = _module(
    "jax.numpy.linalg",
    solve=_linalg_solve, cholesky=_cholesky, eigh=_wrapping(_np.linalg.eigh), svd=_wrapping(_np.linalg.svd),
    qr=_wrapping(_np.linalg.qr), pinv=_wrapping(_np.linalg.pinv), inv=_wrapping(_np.linalg.inv),
    multi_dot=_wrapping(lambda ms: _np.linalg.multi_dot([_np.asarray(m) for m in ms])),
    norm=_wrapping(_np.linalg.norm), det=_wrapping(_np.linalg.det), slogdet=_wrapping(_np.linalg.slogdet),
)

numpy = _module("jax.numpy", linalg=_linalg, pi=_np.pi, inf=_np.inf, nan=_np.nan, newaxis=None,
                float64=_np.float64, float32=_np.float32, int32=_np.int32, int64=_np.int64, uint32=_np.uint32,
                ndarray=_np.ndarray)
for _name in ("array asarray zeros ones eye identity arange linspace concatenate stack hstack vstack sum matmul dot "
              "where reshape repeat meshgrid atleast_1d atleast_2d atleast_3d log exp sqrt sin cos tanh abs maximum "
              "minimum isclose nan_to_num searchsorted nonzero ix_ diag trace outer transpose zeros_like ones_like "
              "full cumsum prod mean var std max min argmax argmin square power isnan isfinite all any allclose "
              "tril triu einsum squeeze expand_dims swapaxes moveaxis tile flip append ravel diagonal kron "
              "log1p expm1 sign clip floor ceil block").split():
    setattr(numpy, _name, _wrapping(getattr(_np, _name)))

# ---------------------------------------------------------------------------------------------------------------------
# jax.scipy
# ---------------------------------------------------------------------------------------------------------------------
_jsl = _module(
    "jax.scipy.linalg",
    block_diag=_wrapping(lambda *ms: _sla.block_diag(*[_np.asarray(m) for m in ms])),
    solve_triangular=_wrapping(lambda a, b, trans=0, lower=False, unit_diagonal=False, **_k:
                               _sla.solve_triangular(_np.asarray(a), _np.asarray(b), trans=trans, lower=lower,
                                                     unit_diagonal=unit_diagonal)),
    cho_factor=lambda a, lower=False, **_k: (_wrap(_sla.cho_factor(_np.asarray(a), lower=lower)[0]), lower),
    cho_solve=_wrapping(lambda c_low, b, **_k: _sla.cho_solve((_np.asarray(c_low[0]), c_low[1]), _np.asarray(b))),
    cholesky=_wrapping(lambda a, lower=False, **_k: _sla.cholesky(_np.asarray(a), lower=lower)),
    expm=_wrapping(_sla.expm),
)
_jss = _module("jax.scipy.special", gammaln=_wrapping(_ssp.gammaln), erf=_wrapping(_ssp.erf),
               logsumexp=_wrapping(_ssp.logsumexp))
def _mvn_logpdf(x, mean, cov, allow_singular=None):
    """jax.scipy.stats.multivariate_normal.logpdf: Cholesky factor, triangular solve, no PSD screening (unlike SciPy's,
    which eigendecomposes and rejects / truncates small eigenvalues)."""
    x, mean, cov = _np.asarray(x, dtype=_np.float64), _np.asarray(mean, dtype=_np.float64), _np.asarray(cov)
    n = mean.shape[-1]
    if cov.ndim < 2:
        return -0.5 * (n * _np.log(2 * _np.pi) + _np.log(cov) + (x - mean) ** 2 / cov)
    L = _np.asarray(_cholesky(cov))
    y = _sla.solve_triangular(L, x - mean, lower=True)
    return -0.5 * _np.sum(y * y) - n / 2.0 * _np.log(2 * _np.pi) - _np.sum(_np.log(_np.diagonal(L)))


_mvn = types.SimpleNamespace(logpdf=_wrapping(_mvn_logpdf))
_norm = types.SimpleNamespace(logpdf=_wrapping(lambda x, loc=0, scale=1: _sst.norm.logpdf(x, loc, scale)))
_jst = _module("jax.scipy.stats", multivariate_normal=_mvn, norm=_norm)
scipy = _module("jax.scipy", linalg=_jsl, special=_jss, stats=_jst, cho_factor=_jsl.cho_factor,
                cho_solve=_jsl.cho_solve)


# ---------------------------------------------------------------------------------------------------------------------
# jax.lax
# ---------------------------------------------------------------------------------------------------------------------
def _scan(f, init, xs=None, length=None, reverse=False, unroll=1):
    if xs is None:
        n = length
    else:
        n = _np.shape(_tree_leaves(xs)[0])[0]
    carry = init
    ys = [None] * n
    order = range(n - 1, -1, -1) if reverse else range(n)
    for i in order:
        x = None if xs is None else _tree_map(lambda a: _wrap(_np.asarray(a)[i]), xs)
        carry, y = f(carry, x)
        ys[i] = y
    if n == 0 or all(y is None for y in ys):
        return carry, None
    return carry, _stack(ys)


def _cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def _fori_loop(lo, hi, body, init):
    v = init
    for i in range(int(lo), int(hi)):
        v = body(i, v)
    return v


lax = _module("jax.lax", scan=_scan, cond=_cond, fori_loop=_fori_loop)


# ---------------------------------------------------------------------------------------------------------------------
# jax.random: keyed, counter-based (NumPy Philox), and LOGGED
# ---------------------------------------------------------------------------------------------------------------------
DRAW_LOG = []          # (key as a 2-tuple, "normal" | "mvn-<method>", z) in call order


def _key_tuple(key):
    k = _np.asarray(key).astype(_np.uint64).ravel()
    return (int(k[0]), int(k[1]))


def _rng(key):
    return _np.random.Generator(_np.random.Philox(key=list(_key_tuple(key))))


def _PRNGKey(seed):
    return Array(_np.array([0, int(seed) & 0xFFFFFFFF], dtype=_np.uint32))


def _split(key, num=2):
    shape = tuple(num) if isinstance(num, (tuple, list)) else (int(num),)
    bits = _rng(key).integers(0, 2 ** 32, size=shape + (2,), dtype=_np.uint64).astype(_np.uint32)
    return Array(bits)


def _normal(key, shape=(), dtype=_np.float64):
    z = _rng(key).standard_normal(tuple(shape))
    DRAW_LOG.append((_key_tuple(key), "normal", z.copy()))
    return Array(z)


def _multivariate_normal(key, mean, cov, shape=None, dtype=None, method="cholesky"):
    """jax.random.multivariate_normal: mean + factor @ z with factor = cholesky(cov) (default), U sqrt(s) from the SVD
    (method='svd') or V sqrt(w) from eigh (method='eigh')."""
    mean, cov = _np.asarray(mean), _np.asarray(cov)
    assert shape is None
    z = _rng(key).standard_normal(mean.shape)
    DRAW_LOG.append((_key_tuple(key), "mvn-" + method, z.copy()))
    if method == "svd":
        u, s, _ = _np.linalg.svd(cov)
        factor = u * _np.sqrt(s[..., None, :])
    elif method == "eigh":
        w, v = _np.linalg.eigh(cov)
        factor = v * _np.sqrt(w[..., None, :])
    else:
        factor = _np.asarray(_cholesky(cov))
    return Array(mean + _np.einsum("...ij,...j->...i", factor, z))


def _randint(key, shape, minval, maxval, dtype=_np.int64):
    return Array(_rng(key).integers(minval, maxval, size=tuple(shape)))


def _uniform(key, shape=(), dtype=_np.float64, minval=0.0, maxval=1.0):
    return Array(_rng(key).uniform(minval, maxval, size=tuple(shape)))


random = _module("jax.random", PRNGKey=_PRNGKey, key=_PRNGKey, split=_split, normal=_normal,
                 multivariate_normal=_multivariate_normal, randint=_randint, uniform=_uniform, DRAW_LOG=DRAW_LOG)

tree_util = _module("jax.tree_util", tree_map=_tree_map, tree_leaves=_tree_leaves)
tree_map = _tree_map
