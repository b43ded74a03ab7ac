"""
The oracle's analytic M-step gradients (oracle.elbo_gradients: dELBO / d kernel variance, lengthscales, Z, likelihood
parameter, sites fixed) against central differences of the reference-order OracleTSVGP.elbo — the role TensorFlow autodiff
plays in the reference (tsvgp.py:79-95; tests/models/test_tsvgp.py:168-188).
"""
import copy

import numpy as np
import pytest

from oracle import tsvgp_oracle as orc


def _model(kind, lik_name, ard):
    rng = np.random.default_rng(3)
    N, M, D = 120, 9, 3
    X = rng.standard_normal((N, D))
    Z = X[:M] + 0.1 * rng.standard_normal((M, D))
    f = np.sin(X.sum(1, keepdims=True))
    ls = np.array([1.3, 1.7, 2.1]) if ard else 1.6
    kernel = (orc.SquaredExponential if kind == "se" else orc.Matern52)(variance=1.4, lengthscales=ls)
    if lik_name == "gaussian":
        lik, Y = orc.Gaussian(variance=0.2), f + 0.3 * rng.standard_normal((N, 1))
    elif lik_name == "bernoulli":
        lik, Y = orc.Bernoulli(), (f + 0.3 * rng.standard_normal((N, 1)) > 0).astype(float)
    else:
        lik, Y = orc.StudentT(scale=0.4, df=3.0), f + 0.3 * rng.standard_t(3.0, size=(N, 1))
    m = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(Z), num_data=4 * N)
    for _ in range(2):
        m.natgrad_step((X, Y), lr=0.6)
    return m, (X, Y)


def _fd(m, data, setter, x0, h):
    def f(x):
        mm = copy.copy(m)
        mm.kernel, mm.likelihood, mm.inducing_variable = copy.deepcopy(m.kernel), copy.deepcopy(m.likelihood), copy.deepcopy(m.inducing_variable)
        setter(mm, x)
        return mm.elbo(data)
    return (f(x0 + h) - f(x0 - h)) / (2 * h)


@pytest.mark.parametrize("kind,lik_name,ard", [("se", "gaussian", False), ("matern52", "gaussian", True), ("se", "bernoulli", True),
                                               ("matern52", "student_t", False)])
def test_analytic_gradients_match_central_differences(kind, lik_name, ard):
    m, data = _model(kind, lik_name, ard)
    elbo, g = orc.elbo_gradients(m, data)
    assert abs(elbo - m.elbo(data)) < 1e-9 * abs(elbo)
    tol = dict(rtol=2e-6, atol=1e-6)
    fd = _fd(m, data, lambda mm, x: setattr(mm.kernel, "variance", orc._param(x)), float(m.kernel.variance), 1e-5)
    np.testing.assert_allclose(g["variance"], fd, **tol)
    ls0 = np.atleast_1d(np.asarray(m.kernel.lengthscales, dtype=float))
    for d in range(ls0.size):
        def set_ls(mm, x, d=d):
            l = ls0.copy(); l[d] = x
            mm.kernel.lengthscales = orc._param(l if ls0.size > 1 else l[0])
        np.testing.assert_allclose(g["lengthscales"][d], _fd(m, data, set_ls, ls0[d], 1e-5), **tol)
    Z0 = np.asarray(m.inducing_variable.Z).copy()
    for (i, d) in [(0, 0), (3, 1), (8, 2)]:
        def set_z(mm, x, i=i, d=d):
            Zn = Z0.copy(); Zn[i, d] = x
            mm.inducing_variable = orc.InducingPoints(Zn)
        np.testing.assert_allclose(g["Z"][i, d], _fd(m, data, set_z, Z0[i, d], 1e-5), **tol)
    if lik_name == "gaussian":
        fd = _fd(m, data, lambda mm, x: setattr(mm.likelihood, "variance", orc._param(x)), float(m.likelihood.variance), 1e-6)
        np.testing.assert_allclose(g["likelihood"], fd, **tol)
    elif lik_name == "student_t":
        fd = _fd(m, data, lambda mm, x: setattr(mm.likelihood, "scale", orc._param(x)), float(m.likelihood.scale), 1e-6)
        np.testing.assert_allclose(g["likelihood"], fd, **tol)
    else:
        assert g["likelihood"] is None
