"""
The reformulated algebra the CUDA path executes (tests/algo_model.py; DESIGN.md "Algebra") against the reference-order
oracle, on the CPU.  This is what makes a GPU mismatch attributable to a kernel rather than to the maths.
"""
import numpy as np
import pytest

from oracle import tsvgp_oracle as orc
from tests import algo_model as am


def relerr(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))


@pytest.mark.parametrize("lik_name", ["gaussian", "bernoulli", "student_t"])
def test_reformulation_matches_reference_order(lik_name):
    rng = np.random.default_rng(5)
    N, M, D = 700, 60, 4
    X = rng.standard_normal((N, D))
    Z = X[:M].copy()
    f = np.sin(X.sum(1, keepdims=True))
    kernel = orc.Matern52(variance=1.3, lengthscales=1.5)
    if lik_name == "gaussian":
        lik, Y = orc.Gaussian(variance=0.1), f + 0.3 * rng.standard_normal((N, 1))
    elif lik_name == "bernoulli":
        lik, Y = orc.Bernoulli(), (f + 0.3 * rng.standard_normal((N, 1)) > 0).astype(float)
    else:
        lik, Y = orc.StudentT(scale=0.3, df=3.0), f + 0.3 * rng.standard_t(3.0, size=(N, 1))
    ref = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(Z), num_data=5 * N)
    lr, scale = 0.6, 5.0
    ref.natgrad_step((X, Y), lr=lr)   # non-trivial sites
    l1, L2 = ref.lambda_1.copy(), ref.lambda_2_sqrt[0].copy()

    K = kernel.K(Z)
    Kuf = kernel.K(Z, X)
    pre = am.prepare(K, l1, L2)
    mu, var = am.marginals(Kuf, kernel.K_diag(X), pre["T"], pre["alpha"])
    mu = mu[:, 0]
    mu_r, var_r = ref.predict_f(X)
    assert relerr(mu, mu_r[:, 0]) < 1e-10 and relerr(var, var_r[:, 0]) < 1e-10
    ve, g, h = lik.ve_and_grads(mu[:, None], var[:, None], Y)
    h = np.minimum(h, -1e-8)
    elbo = scale * np.sum(ve) - am.kl(pre["K6"], pre["T"], pre["alpha"], pre["Uw"])
    assert abs(elbo - ref.elbo((X, Y))) < 1e-10 * abs(elbo)
    for whiten in (False, True):
        n1, nL2 = am.natgrad(K, Kuf, g[:, 0], h[:, 0], pre["alpha"], l1, L2, lr, scale, whiten=whiten)
        ref2 = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(Z), num_data=5 * N, lambda_1=l1, lambda_2_sqrt=L2[None])
        ref2.natgrad_step((X, Y), lr=lr)
        assert relerr(n1, ref2.lambda_1) < 1e-9
        assert relerr(nL2 @ nL2.T, ref2.lambda_2[0]) < 1e-9
