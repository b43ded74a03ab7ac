"""
1-D regression with variational EM on the B200 path — the counterpart of the reference's docs/notebooks/regression_1D.py
(same toy data and model sizes) followed by hyperparameter learning as in experiments/uci_regression.py.
With GPflow installed, pass gpflow kernel / likelihood objects instead of the attribute-only stand-ins: only attributes are read.

    python examples/vem_regression_1d.py        (needs a B200; there is no CPU fallback)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tsvgp_b200 as tb  # noqa: E402
from tsvgp_b200 import standins as st, vem  # noqa: E402


def main(n_iters=15, verbose=True):
    rng = np.random.RandomState(123)
    N, M = 200, 20
    X = rng.rand(N, 1) * 2 - 1
    Y = np.sin(15 * X) + 0.3 * rng.randn(N, 1)               # true noise variance 0.09
    Z = np.linspace(X.min(), X.max(), M).reshape(-1, 1)
    kernel = st.SquaredExponential(variance=0.3, lengthscales=0.3)   # deliberately poor initial hyperparameters
    lik = st.Gaussian(variance=1.0)
    model = tb.t_SVGP(kernel, lik, Z, num_data=N)
    for _ in range(5):                                       # regression_1D.py: nit = 5 natural-gradient steps, lr 0.9
        model.natgrad_step((X, Y), lr=0.9)
    e0 = model.elbo((X, Y))
    trace = vem.fit(model, (X, Y), n_iters=n_iters, n_e_steps=4, n_m_steps=10, lr_natgrad=0.9, lr_adam=0.05,
                    callback=(lambda it, e, m: print(f"iter {it:2d}  elbo {e:10.4f}  variance {m.kernel.variance:.3f}  "
                                                     f"lengthscale {float(m.kernel.lengthscales):.3f}  noise {m.likelihood.variance:.3f}"))
                    if verbose else None)
    for _ in range(4):
        model.natgrad_step(lr=0.9)
    e1 = model.elbo()
    xg = np.linspace(-1, 1, 100)[:, None]
    mu, var = model.predict_y(xg)
    rmse = float(np.sqrt(np.mean((mu[:, 0] - np.sin(15 * xg[:, 0])) ** 2)))
    if verbose:
        print(f"ELBO {e0:.3f} -> {e1:.3f}; learned noise variance {lik.variance:.3f} (true 0.09); RMSE to the true function {rmse:.3f}")
    model.close()
    return e0, e1, float(lik.variance), rmse, trace


if __name__ == "__main__":
    main()
