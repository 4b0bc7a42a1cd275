"""ODE right-hand sides known to the device library.

The reference takes ``ode_fun(X, t, **params)`` as a traced Python callable (src/rodeo/solve.py:219).  A Python
callable cannot run inside a CUDA kernel, so here ``ode_fun`` is an :class:`OdeModel`: a handle to a
``__device__`` functor, either compiled into librodeo_b200.so (rodeo_b200/csrc/rodeo_models.cuh) or given as a CUDA
source string and compiled by NVRTC (:class:`CudaOde`).  Passing any other callable raises ``NotImplementedError``
-- there is no CPU fallback.
"""


class OdeModel:
    def __init__(self, name, model_id, n_block, n_bstate, n_bmeas, n_theta, doc="", wcol=1):
        self.name, self.model_id = name, model_id
        self.n_block, self.n_bstate, self.n_bmeas, self.n_theta = n_block, n_bstate, n_bmeas, n_theta
        self.wcol = wcol            # the ODE is X[:, wcol] = f(X, t): W = e_wcol
        self.__doc__ = doc

    def __repr__(self):
        return f"<rodeo_b200.OdeModel {self.name} n_block={self.n_block} n_bstate={self.n_bstate}>"

    def __call__(self, *a, **k):
        raise NotImplementedError(
            f"{self.name} is a device functor; it is evaluated inside the CUDA kernels "
            "(use rodeo_b200.utils.first_order_pad(...)[1] to evaluate it for initial values)")


fitzhugh_nagumo = OdeModel("fitzhugh_nagumo", 0, 2, 3, 1, 3,
                           "FitzHugh-Nagumo, theta=(a,b,c) (reference README.md:92-99)")
lorenz63 = OdeModel("lorenz63", 1, 3, 3, 1, 3,
                    "Lorenz63, theta=(rho,sigma,beta) (reference docs/examples/lorenz.md:95-101)")
second_order_sin = OdeModel("second_order_sin", 2, 1, 4, 1, 2,
                            "x''=sin(omega t)-k x, theta=(omega,k) (reference docs/examples/higher_order.md:47-58)",
                            wcol=2)
hes1 = OdeModel("hes1", 3, 3, 3, 1, 7, "log-scale Hes1 (reference examples/timings.py:253-262)")
seirah = OdeModel("seirah", 4, 6, 3, 1, 6, "SEIRAH (reference examples/timings.py:339-351)")
pair_one_block = OdeModel("pair_one_block", 5, 1, 6, 2, 3,
                          "n_bmeas = 2: (x, x', x'', y, y', y'') in one block, x' = -a x + sin t, y' = -b y^2, "
                          "theta=(a,b,c); float64 solve_mv / dalton / fenrir (tests/golden/make_reference_golden.py)")

BUILTIN = {m.name: m for m in (fitzhugh_nagumo, lorenz63, second_order_sin, hes1, seirah, pair_one_block)}

_USER_TEMPLATE = """
struct UserModel {{
  static constexpr int NB = {nb}, P = {p}, M = 1, NTHETA = {nth}, JCOLS = {jc}, WCOL = {wcol};
  static constexpr bool USES_TIME = true, HAS_JAC = false;       // Jacobian by dual numbers (== jax.jacfwd)
  template <class T> struct Par {{ T th[{nth1}]; }};
  template <class T> RD_DEV static Par<T> load(const T* p) {{
    Par<T> q;
    for (int k = 0; k < {nth}; ++k) q.th[k] = p[k];
    return q;
  }}
  template <class T, class X>
  RD_DEV static void rhs(const Par<T>& q, T t, const X (&x)[NB][JCOLS], X (&f)[NB][M]) {{
    using namespace rodeo;
    const T* th = q.th;
    (void)th; (void)t;
    {body}
  }}
  template <class T>
  RD_DEV static void jac(const Par<T>&, T, const T (&)[NB][JCOLS], T (&)[NB][M][JCOLS]) {{}}
}};
"""


class CudaOde(OdeModel):
    """A user ODE right-hand side as a CUDA source string, compiled for sm_100a by NVRTC on first use.

    The reference accepts any Python callable ``ode_fun(X, t, **params)``; the device-side equivalent is a few lines
    of C++ assigning ``f[b][0]`` for every block ``b`` from ``x[b][j]`` (state column ``j < jcols`` of block ``b``),
    the time ``t`` and the parameters ``th[k]``.  Write it against the value type ``X`` and the scalar type ``T`` so
    that the same code runs on dual numbers: ``interrogate_kramer``'s block-diagonal Jacobian is obtained by
    forward-mode differentiation of this very function, exactly like ``jax.jacfwd`` in the reference::

        fitz = CudaOde("my_fitz", n_block=2, n_bstate=3, n_theta=3, rhs=(
            "X V = x[0][0], R = x[1][0];"
            "f[0][0] = th[2] * (V - V * V * V / T(3) + R);"
            "f[1][0] = T(-1) / th[2] * (V - th[0] + th[1] * R);"))

    ``wcol``: the ODE is ``X[:, wcol] = f(X, t)`` (1 for a first-order system padded by ``first_order_pad``).
    ``source=`` may instead give a complete ``struct UserModel`` (see rodeo_b200/csrc/rodeo_models.cuh).
    """

    def __init__(self, name, n_block, n_bstate, n_theta, rhs=None, jcols=1, wcol=1, source=None):
        import ctypes
        from . import _lib
        if (rhs is None) == (source is None):
            raise TypeError("give exactly one of rhs= (function body) or source= (full struct UserModel)")
        if source is None:
            source = _USER_TEMPLATE.format(nb=n_block, p=n_bstate, nth=n_theta, nth1=max(n_theta, 1), jc=jcols,
                                           wcol=wcol, body=rhs)
        self.source = source
        lib = _lib.load()
        mid = ctypes.c_int(0)
        rc = lib.rodeo_b200_register_model_nvrtc(name.encode(), source.encode(), n_block, n_bstate, 1, n_theta,
                                                 ctypes.byref(mid))
        _lib.check(rc, "register_model_nvrtc")
        super().__init__(name, mid.value, n_block, n_bstate, 1, n_theta, doc="user CUDA model", wcol=wcol)


def resolve(ode_fun):
    if isinstance(ode_fun, OdeModel):
        return ode_fun
    if isinstance(ode_fun, str) and ode_fun in BUILTIN:
        return BUILTIN[ode_fun]
    raise NotImplementedError(
        "ode_fun must be a rodeo_b200.models.OdeModel (a built-in device functor or a CudaOde source string); "
        "arbitrary Python callables cannot be traced into the CUDA kernels and there is no CPU fallback")
