"""ODE right-hand sides known to the device library.

The reference takes ``ode_fun(X, t, **params)`` as a traced Python callable (src/rodeo/solve.py:219).  A Python
callable cannot run inside a CUDA kernel, so here ``ode_fun`` is an :class:`OdeModel`: a handle to a
``__device__`` functor compiled into librodeo_b200.so (rodeo_b200/csrc/rodeo_models.cuh).  Passing any other
callable raises ``NotImplementedError`` -- there is no CPU fallback.
"""


class OdeModel:
    def __init__(self, name, model_id, n_block, n_bstate, n_bmeas, n_theta, doc=""):
        self.name, self.model_id = name, model_id
        self.n_block, self.n_bstate, self.n_bmeas, self.n_theta = n_block, n_bstate, n_bmeas, n_theta
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
                            "x''=sin(omega t)-k x, theta=(omega,k) (reference docs/examples/higher_order.md:47-58)")
hes1 = OdeModel("hes1", 3, 3, 3, 1, 7, "log-scale Hes1 (reference examples/timings.py:253-262)")
seirah = OdeModel("seirah", 4, 6, 3, 1, 6, "SEIRAH (reference examples/timings.py:339-351)")

BUILTIN = {m.name: m for m in (fitzhugh_nagumo, lorenz63, second_order_sin, hes1, seirah)}


def resolve(ode_fun):
    if isinstance(ode_fun, OdeModel):
        return ode_fun
    if isinstance(ode_fun, str) and ode_fun in BUILTIN:
        return BUILTIN[ode_fun]
    raise NotImplementedError(
        "ode_fun must be a rodeo_b200.models.OdeModel (a device functor); arbitrary Python callables cannot be "
        "traced into the CUDA kernels and there is no CPU fallback")
