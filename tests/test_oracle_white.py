"""
Pins the oracle's restatement of the whitened sibling `t_SVGP_white` by the reference's own property tests
(reference tests/models/test_tsvgp_white.py:36-85, 88-110, all decimal=4) and tests/test_utils.py:124-137
(posterior_from_dense_site vs posterior_from_dense_site_white).
"""
import numpy as np

from oracle import tsvgp_oracle as orc


def _setup():
    rng = np.random.RandomState(123)
    X = rng.rand(8, 1) * 2 - 1
    Y = np.sin(X * 3 * 3.14) + 0.3 * np.cos(X * 9 * 3.14) + 0.5 * np.sin(X * 7 * 3.14) + 0.2 * rng.randn(8, 1)
    return X, Y, orc.SquaredExponential(lengthscales=2.0, variance=2.25), 0.3


def test_white_and_plain_agree_at_init_and_after_one_step():   # test_tsvgp_white.py:58-85
    X, Y, kernel, s2 = _setup()
    a = orc.OracleTSVGP(kernel, orc.Gaussian(variance=s2), orc.InducingPoints(X.copy()))
    b = orc.OracleTSVGPWhite(kernel, orc.Gaussian(variance=s2), orc.InducingPoints(X.copy()))
    np.testing.assert_almost_equal(a.elbo((X, Y)), b.elbo((X, Y)), decimal=4)
    for m in (a, b):
        m.natgrad_step((X, Y), lr=0.9)
    (ma, va), (mb, vb) = a.predict_f(X), b.predict_f(X)
    np.testing.assert_array_almost_equal(ma, mb, decimal=4)
    np.testing.assert_array_almost_equal(va, vb, decimal=4)


def test_one_full_step_of_the_white_model_is_gp_regression():   # fixture of test_tsvgp_white.py:88-110
    X, Y, kernel, s2 = _setup()
    Y = Y * 0                                                   # as the reference's fixture (:93)
    m = orc.OracleTSVGPWhite(kernel, orc.Gaussian(variance=s2), orc.InducingPoints(X.copy()))
    m.natgrad_step((X, Y), lr=1.0)
    np.testing.assert_almost_equal(m.elbo((X, Y)), orc.gpr_log_marginal_likelihood(kernel, X, Y, s2), decimal=3)
    mu, var = m.predict_f(X + 1.0)
    mu_g, var_g = orc.gpr_predict_f(kernel, X, Y, s2, X + 1.0)
    np.testing.assert_array_almost_equal(mu, mu_g, decimal=4)
    np.testing.assert_array_almost_equal(var, var_g, decimal=4)


def test_posteriors_of_the_two_parameterisations_match():       # reference tests/test_utils.py:124-137 (decimal=3)
    rng = np.random.RandomState(123)
    Z = rng.rand(5, 1) * 2 - 1
    K = orc.SquaredExponential(lengthscales=0.7, variance=2.25).K(Z) + 1e-6 * np.eye(5)
    L = np.tril(rng.randn(1, 5, 5))
    l1 = rng.randn(5, 1)
    m1, cs1 = orc.posterior_from_dense_site(K, l1, L)
    m2, cs2 = orc.posterior_from_dense_site_white(K, K @ l1, (K @ L[0] @ L[0].T @ K)[None])
    np.testing.assert_array_almost_equal(m1, m2, decimal=3)
    np.testing.assert_array_almost_equal(cs1[0] @ cs1[0].T, cs2[0] @ cs2[0].T, decimal=3)


def test_white_model_with_several_latents_is_independent_single_latent_models():
    # tsvgp_white.py:79-89, util.py:60-88 / :264-291 / :411-426: the latents share the kernel and Z and nothing else, so the L-latent
    # restatement (what tests/test_gpu_white.py::test_white_model_with_several_latents holds the device to) must equal L single-latent
    # models run side by side: sites, predictions, posterior, predict_f_extra_data; ELBO and KL are the sums.
    rng = np.random.RandomState(5)
    n, M, L = 60, 12, 3
    X = rng.rand(n, 2) * 2 - 1
    Z = X[:M].copy()
    Y = np.stack([np.sin(3 * X[:, 0] + l) + 0.1 * rng.randn(n) for l in range(L)], 1)
    kernel, lik = orc.SquaredExponential(lengthscales=0.9, variance=1.5), orc.Gaussian(variance=0.2)
    multi = orc.OracleTSVGPWhite(kernel, lik, orc.InducingPoints(Z.copy()), num_latent_gps=L, num_data=3 * n)
    singles = [orc.OracleTSVGPWhite(kernel, lik, orc.InducingPoints(Z.copy()), num_data=3 * n) for _ in range(L)]
    for _ in range(2):
        e_multi = multi.elbo((X, Y))
        e_single = sum(s.elbo((X, Y[:, l:l + 1])) for l, s in enumerate(singles))
        np.testing.assert_allclose(e_multi, e_single, rtol=1e-11)
        multi.natgrad_step((X, Y), lr=0.7)
        for l, s in enumerate(singles):
            s.natgrad_step((X, Y[:, l:l + 1]), lr=0.7)
    assert multi.lambda_1.shape == (M, L) and multi.lambda_2.shape == (L, M, M)
    mu, var = multi.predict_f(X[:20] + 0.05)
    m_q, chol_S = multi.get_mean_chol_cov_inducing_posterior()
    mu_e, var_e = multi.predict_f_extra_data(X[:20] + 0.05, (X[:30] * 0.9, Y[:30]))
    for l, s in enumerate(singles):
        np.testing.assert_allclose(multi.lambda_1[:, l], s.lambda_1[:, 0], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(multi.lambda_2[l], s.lambda_2[0], rtol=1e-10, atol=1e-10)
        mu_s, var_s = s.predict_f(X[:20] + 0.05)
        np.testing.assert_allclose(mu[:, l], mu_s[:, 0], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(var[:, l], var_s[:, 0], rtol=1e-9, atol=1e-11)
        m_s, cs_s = s.get_mean_chol_cov_inducing_posterior()
        np.testing.assert_allclose(m_q[:, l], m_s[:, 0], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(chol_S[l] @ chol_S[l].T, cs_s[0] @ cs_s[0].T, rtol=1e-9, atol=1e-11)
        mu_es, var_es = s.predict_f_extra_data(X[:20] + 0.05, (X[:30] * 0.9, Y[:30, l:l + 1]))
        np.testing.assert_allclose(mu_e[:, l], mu_es[:, 0], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(var_e[:, l], var_es[:, 0], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(multi.prior_kl(), sum(s.prior_kl() for s in singles), rtol=1e-10)
