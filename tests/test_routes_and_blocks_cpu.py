"""
CPU checks of two design decisions of the CUDA path, on the NumPy model of its algebra (tests/algo_model.py):
  * the three statistics routes agree on well-conditioned inputs, and at extreme conditioning only the exact route (the
    reference's own order, tsvgp.py:271-281: A = K9^-1 Kuf first, then the Gram product) keeps -2 Lambda_2 + jitter I positive
    definite — which is why the device selects it there (GPU counterpart:
    tests/test_gpu_parity.py::test_extreme_conditioning_keeps_the_reference_error_behaviour);
  * the block schedule of diag_potrf_inv_blocked_kernel (inverse riding on the factorisation, overwrite-on-first-touch of the
    inverse's off-diagonal blocks) is a Cholesky factorisation and its inverse.
"""
import numpy as np

from oracle import tsvgp_oracle as orc
from tests import algo_model as am


def _example_data():   # examples/c_api_example.c
    N, M, s = 2000, 40, 12345
    X, Y = np.zeros((N, 1)), np.zeros((N, 1))
    for i in range(N):
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        X[i] = 2.0 * (s >> 8) / 16777216.0 - 1.0
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        Y[i] = np.sin(6.0 * X[i]) + 0.2 * ((s >> 8) / 16777216.0 - 0.5)
    return X, Y, np.linspace(-1.0, 1.0, M)[:, None]


def _min_eig_of_update(G2, lr=0.9, jitter=1e-9):
    G2 = np.tril(G2) + np.tril(G2, -1).T          # the device reads the lower triangle
    return float(np.linalg.eigvalsh(-2.0 * lr * G2 + jitter * np.eye(G2.shape[0])).min())


def test_routes_agree_when_well_conditioned():
    rng = np.random.default_rng(3)
    X = rng.standard_normal((500, 3))
    Z = X[:40].copy()
    k = orc.Matern52(variance=1.2, lengthscales=1.0)
    K, Kuf = k.K(Z), k.K(Z, X)
    g, h = rng.standard_normal(500), -np.abs(rng.standard_normal(500)) - 0.1
    ref = am.natural_gradients(K, Kuf, g, h, route="exact")
    for route in ("fused", "whitened"):
        G1, G2 = am.natural_gradients(K, Kuf, g, h, route=route)
        assert np.max(np.abs(G1 - ref[0])) <= 1e-9 * np.max(np.abs(ref[0]))
        assert np.max(np.abs(G2 - ref[1])) <= 1e-9 * np.max(np.abs(ref[1]))


def test_only_the_exact_route_survives_extreme_conditioning():
    X, Y, Z = _example_data()
    k = orc.SquaredExponential(variance=1.0, lengthscales=0.2)
    K, Kuf = k.K(Z), k.K(Z, X)
    assert np.linalg.cond(K) > 1e16 and np.linalg.cond(K + 1e-9 * np.eye(40)) > 1e8   # above route_exact_min
    h = np.full(X.shape[0], -0.5 / 0.05)                                              # Gaussian likelihood, variance 0.05
    g = np.zeros(X.shape[0])
    assert _min_eig_of_update(am.natural_gradients(K, Kuf, g, h, route="exact")[1]) > 0.0
    assert _min_eig_of_update(am.natural_gradients(K, Kuf, g, h, route="whitened")[1]) < 0.0
    assert _min_eig_of_update(am.natural_gradients(K, Kuf, g, h, route="fused")[1]) < 0.0
    # and the reference-order oracle indeed takes three steps on this data without raising
    m = orc.OracleTSVGP(k, orc.Gaussian(variance=0.05), orc.InducingPoints(Z.copy()))
    for _ in range(3):
        m.natgrad_step((X, Y), lr=0.9)
    assert np.isfinite(m.elbo((X, Y)))


def test_block_schedule_of_the_diagonal_block_kernel():
    rng = np.random.default_rng(0)
    G = rng.standard_normal((128, 128))
    A = G @ G.T + 128.0 * np.eye(128)
    L, X = am.blocked_potrf_inv(A)
    assert np.max(np.abs(L - np.linalg.cholesky(A))) < 1e-12
    assert np.max(np.abs(X @ L - np.eye(128))) < 1e-13
    assert np.allclose(np.triu(L, 1), 0.0) and np.allclose(np.triu(X, 1), 0.0)
