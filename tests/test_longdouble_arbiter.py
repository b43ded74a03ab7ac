"""
The float64 oracle against the extended-precision arbiter (oracle/longdouble.py) in the same operation order: pins the
oracle's own rounding level (eps * cond) so that GPU-vs-oracle tolerances are read against it.
"""
import numpy as np
import pytest

from oracle import longdouble as ld
from oracle import tsvgp_oracle as orc


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


@pytest.mark.parametrize("lengthscale,bound", [(1.0, 1e-13), (4.0, 1e-11), (16.0, 1e-8)])
def test_oracle_rounding_level_tracks_conditioning(lengthscale, bound):
    rng = np.random.RandomState(0)
    N, M, D = 150, 24, 8
    X, Z = rng.randn(N, D), rng.randn(M, D)
    Y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.randn(N, 1)
    kernel, lik = orc.SquaredExponential(variance=1.0, lengthscales=lengthscale), orc.Gaussian(variance=0.1)
    m = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(Z))
    m.natgrad_step((X, Y), lr=0.7)                                    # non-trivial sites in float64
    l1, L2 = m.lambda_1[:, 0].copy(), m.lambda_2_sqrt[0].copy()
    t1, tL2 = ld.natgrad_step_gaussian(X, Y, Z, 1.0, lengthscale, 0.1, l1, L2, lr=0.7)
    m.natgrad_step((X, Y), lr=0.7)
    cond = np.linalg.cond(kernel.K(Z) + 1e-9 * np.eye(M))
    e1, e2 = relerr(m.lambda_1[:, 0], t1), relerr(m.lambda_2[0], tL2 @ tL2.T)
    assert e1 < bound and e2 < bound, (cond, e1, e2)
    assert max(e1, e2) < 1e-14 * cond + 1e-14                          # ~ eps * cond(Kuu + jitter I)


@pytest.mark.parametrize("lik_name", ["bernoulli", "student_t"])
@pytest.mark.parametrize("lengthscale,bound", [(1.0, 1e-12), (4.0, 1e-10)])
def test_quadrature_path_against_the_long_double_arbiter(lik_name, lengthscale, bound):
    # the Gauss-Hermite path (tsvgp.py:256-263): quadrature sums, analytic gradients, the -1e-8 clip — float64 oracle vs the same
    # formulas in 80-bit arithmetic (erf through mpmath)
    rng = np.random.RandomState(1)
    N, M, D = 120, 20, 6
    X, Z = rng.randn(N, D), rng.randn(M, D)
    f = np.sin(X.sum(1, keepdims=True))
    kernel = orc.SquaredExponential(variance=1.0, lengthscales=lengthscale)
    if lik_name == "bernoulli":
        lik, spec, Y = orc.Bernoulli(), ("bernoulli",), (f + 0.3 * rng.randn(N, 1) > 0).astype(float)
    else:
        lik, spec, Y = orc.StudentT(scale=0.3, df=3.0), ("student_t", 0.3, 3.0), f + 0.3 * rng.standard_t(3.0, size=(N, 1))
    m = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(Z), num_data=4 * N)
    m.natgrad_step((X, Y), lr=0.6)
    l1, L2 = m.lambda_1[:, 0].copy(), m.lambda_2_sqrt[0].copy()
    t1, tL2 = ld.natgrad_step(X, Y, Z, 1.0, lengthscale, spec, l1, L2, lr=0.6, scale=4.0)
    m.natgrad_step((X, Y), lr=0.6)
    e1, e2 = relerr(m.lambda_1[:, 0], t1), relerr(m.lambda_2[0], tL2 @ tL2.T)
    assert e1 < bound and e2 < bound, (e1, e2)
