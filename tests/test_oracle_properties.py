"""
Pins the NumPy oracle by the reference's own property tests (all `decimal=4` there):
reference tests/models/test_tsvgp.py:91-165 and tests/test_utils.py:124-137, restated
against closed-form exact GP regression (what gpflow.models.GPR computes).
"""
import numpy as np
import pytest

from oracle import tsvgp_oracle as orc

LENGTH_SCALE = 2.0
VARIANCE = 2.25
NUM_DATA = 8
NOISE_VARIANCE = 0.3


def _setup(rng):
    # reference tests/models/test_tsvgp.py:91-103
    def func(x):
        return np.sin(x * 3 * 3.14) + 0.3 * np.cos(x * 9 * 3.14) + 0.5 * np.sin(x * 7 * 3.14)

    X = rng.rand(NUM_DATA, 1) * 2 - 1
    Y = func(X) + 0.2 * rng.randn(NUM_DATA, 1)
    kernel = orc.SquaredExponential(lengthscales=LENGTH_SCALE, variance=VARIANCE)
    return X, Y, kernel, NOISE_VARIANCE


@pytest.fixture
def gpr_optim():
    # reference tests/models/test_tsvgp.py:20-42 : 10 natgrad steps lr=0.9, Z = X
    rng = np.random.RandomState(123)
    X, Y, kernel, s2 = _setup(rng)
    m = orc.OracleTSVGP(kernel, orc.Gaussian(variance=s2), orc.InducingPoints(X.copy()))
    for _ in range(10):
        m.natgrad_step((X, Y), lr=0.9)
    return m, (X, Y), kernel, s2


def test_elbo_optimal_equals_gpr_lml(gpr_optim):  # test_tsvgp.py:106-110
    m, (X, Y), kernel, s2 = gpr_optim
    np.testing.assert_almost_equal(m.elbo((X, Y)), orc.gpr_log_marginal_likelihood(kernel, X, Y, s2), decimal=4)


def test_predictions_match_gpr(gpr_optim):  # test_tsvgp.py:113-120
    m, (X, Y), kernel, s2 = gpr_optim
    Xs = X + 1.0
    mu, var = m.predict_f(Xs)
    mu_g, var_g = orc.gpr_predict_f(kernel, X, Y, s2, Xs)
    np.testing.assert_array_almost_equal(mu, mu_g, decimal=4)
    np.testing.assert_array_almost_equal(var, var_g, decimal=4)


def test_unchanged_at_optimum(gpr_optim):  # test_tsvgp.py:134-145
    m, data, _, _ = gpr_optim
    e0 = m.elbo(data)
    m.natgrad_step(data, lr=0.9)
    np.testing.assert_almost_equal(e0, m.elbo(data), decimal=4)


def test_minibatch_same_elbo(gpr_optim):  # test_tsvgp.py:148-165
    m, (X, Y), _, _ = gpr_optim
    x = X[0].repeat(NUM_DATA)[:, None]
    y = Y[0].repeat(NUM_DATA)[:, None]
    e2 = m.elbo((x, y))
    m.num_data = NUM_DATA
    e1 = m.elbo((X[0][:, None], Y[0][:, None]))
    np.testing.assert_almost_equal(e2, e1, decimal=4)


def test_new_predict_f_matches_predict_f(gpr_optim):  # the two algebras of tsvgp.py:97-114 and :215-232 agree
    m, (X, _), _, _ = gpr_optim
    mu1, v1 = m.predict_f(X + 0.1)
    mu2, v2 = m.new_predict_f(X + 0.1)
    np.testing.assert_array_almost_equal(mu1, mu2, decimal=4)
    np.testing.assert_array_almost_equal(v1, v2, decimal=4)


@pytest.mark.parametrize("lik", ["bernoulli", "student_t"])
def test_quadrature_fixed_point_and_gradients(lik):
    # Bernoulli setup of test_tsvgp.py:45-88 (20 steps lr=1.0): natgrad fixed point is reached, and the analytic
    # Gauss-Hermite gradients (replacing tf.GradientTape, tsvgp.py:256-259) match central differences.
    rng = np.random.RandomState(123)
    X, Y, kernel, _ = _setup(rng)
    if lik == "bernoulli":
        Yl, likelihood = (Y > 0).astype(float), orc.Bernoulli()
    else:
        Yl, likelihood = Y, orc.StudentT(scale=0.4, df=3.0)
    m = orc.OracleTSVGP(kernel, likelihood, orc.InducingPoints(X.copy()))
    for _ in range(20):
        m.natgrad_step((X, Yl), lr=1.0)
    l1, l2 = m.lambda_1.copy(), m.lambda_2.copy()
    m.natgrad_step((X, Yl), lr=1.0)
    assert np.max(np.abs(m.lambda_1 - l1)) <= 1e-6 * np.max(np.abs(l1))
    assert np.max(np.abs(m.lambda_2 - l2)) <= 1e-6 * np.max(np.abs(l2))
    mu, var = m.predict_f(X)
    ve, gm, gv = likelihood.ve_and_grads(mu, var, Yl)
    np.testing.assert_allclose(ve, likelihood.variational_expectations(mu, var, Yl), rtol=1e-14)
    h = 1e-6
    gm_fd = (likelihood.variational_expectations(mu + h, var, Yl) - likelihood.variational_expectations(mu - h, var, Yl)) / (2 * h)
    gv_fd = (likelihood.variational_expectations(mu, var + h * var, Yl) - likelihood.variational_expectations(mu, var - h * var, Yl)) / (2 * h * var[:, 0])
    np.testing.assert_allclose(gm[:, 0], gm_fd, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(gv[:, 0], gv_fd, rtol=1e-5, atol=1e-8)


def test_posterior_from_dense_site_is_the_site_posterior():
    # reference tests/test_utils.py:124-137 pins posterior_from_dense_site against the whitened algebra;
    # restated: S = (K^-1 + L L^T)^-1, m = S lambda_1 by direct inversion.
    rng = np.random.RandomState(123)
    Z = rng.rand(5, 1) * 2 - 1
    K = orc.SquaredExponential(lengthscales=0.7, variance=VARIANCE).K(Z) + 1e-6 * np.eye(5)
    L = np.tril(rng.randn(1, 5, 5))
    l1 = rng.randn(5, 1)
    m_q, chol_S = orc.posterior_from_dense_site(K, l1, L)
    S = np.linalg.inv(np.linalg.inv(K) + L[0] @ L[0].T)
    np.testing.assert_array_almost_equal(chol_S[0] @ chol_S[0].T, S, decimal=6)
    np.testing.assert_array_almost_equal(m_q, S @ l1, decimal=6)


def test_gauss_kl_closed_form():
    rng = np.random.RandomState(0)
    M = 6
    A = rng.randn(M, M)
    K = A @ A.T + M * np.eye(M)
    Lq = np.tril(rng.randn(1, M, M)) + 2 * np.eye(M)
    mu = rng.randn(M, 1)
    S = Lq[0] @ Lq[0].T
    kl = 0.5 * (np.trace(np.linalg.solve(K, S)) + mu[:, 0] @ np.linalg.solve(K, mu[:, 0]) - M
                + np.linalg.slogdet(K)[1] - np.linalg.slogdet(S)[1])
    np.testing.assert_allclose(orc.gauss_kl(mu, Lq, K), kl, rtol=1e-12)
