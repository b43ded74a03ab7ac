"""
GPU: num_latent_gps = L > 1 in ONE device context (shared kernel, inducing inputs, Kuf slab; per-latent sites, variance product,
weighted SYRK and site update) against the reference-order oracle — reference src/models/tsvgp.py:243-254,276-281, its L = 2
Bernoulli fixture (tests/models/test_tsvgp.py:45-88) and the Softmax classifier of docs/notebooks/mnist.py:117-122.
Tolerance 1e-9 (norm-wise) on lambda_1 [M, L], lambda_2 [L, M, M], ELBO, predictive moments [N, L].
"""
import numpy as np
import pytest

from oracle import tsvgp_oracle as orc
from tests.test_gpu_parity import check, relerr

pytestmark = pytest.mark.gpu


def _pair_errors(dev, ref, data, Xt, steps, lr):
    errs = {}
    for s in range(steps):
        e_ref = ref.elbo(data)
        e_dev = dev.natgrad_step(data, lr=lr, return_elbo=True)
        ref.natgrad_step(data, lr=lr)
        errs[f"elbo_before_step{s}"] = abs(e_dev - e_ref) / abs(e_ref)
        errs[f"lambda_1_step{s}"] = relerr(dev.lambda_1, ref.lambda_1)
        errs[f"lambda_2_step{s}"] = relerr(dev.lambda_2, ref.lambda_2)
    errs["elbo"] = abs(dev.elbo(data) - ref.elbo(data)) / abs(ref.elbo(data))
    mu_d, var_d = dev.predict_f(Xt)
    mu_r, var_r = ref.predict_f(Xt)
    assert mu_d.shape == mu_r.shape and var_d.shape == var_r.shape
    errs["mean"], errs["var"] = relerr(mu_d, mu_r), relerr(var_d, var_r)
    errs["prior_kl"] = abs(dev.prior_kl() - ref.prior_kl()) / abs(ref.prior_kl())
    m_d, cs_d = dev.get_mean_chol_cov_inducing_posterior()
    m_r, cs_r = ref.get_mean_chol_cov_inducing_posterior()
    errs["m_q"] = relerr(m_d, m_r)
    errs["S_q"] = relerr(cs_d @ np.swapaxes(cs_d, -1, -2), cs_r @ np.swapaxes(cs_r, -1, -2))
    return errs


@pytest.mark.parametrize("lik_name,L,M,N", [("gaussian", 3, 200, 1500), ("bernoulli", 2, 130, 1000), ("student_t", 4, 257, 2100)])
def test_independent_likelihood_terms_over_L_latents(lik_name, L, M, N):
    import tsvgp_b200 as tb
    rng = np.random.default_rng(11)
    D = 5
    X = rng.standard_normal((N, D))
    Z = X[:M].copy()
    F = np.stack([np.sin(X.sum(1) * (0.5 + 0.3 * l)) for l in range(L)], axis=1)
    kernel = orc.Matern52(variance=1.2, lengthscales=1.8)
    if lik_name == "gaussian":
        lik, Y = orc.Gaussian(variance=0.1), F + 0.3 * rng.standard_normal((N, L))
    elif lik_name == "bernoulli":
        lik, Y = orc.Bernoulli(), (F + 0.3 * rng.standard_normal((N, L)) > 0).astype(float)
    else:
        lik, Y = orc.StudentT(scale=0.3, df=3.0), F + 0.3 * rng.standard_t(3.0, size=(N, L))
    ref = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(Z.copy()), num_latent_gps=L, num_data=4 * N)
    dev = tb.t_SVGP(kernel, lik, Z.copy(), num_latent_gps=L, num_data=4 * N)
    dev.set_option("chunk", 512)     # several slabs over two streams
    assert dev.lambda_1.shape == (M, L) and dev.lambda_2_sqrt.shape == (L, M, M)
    check(_pair_errors(dev, ref, (X, Y), X[:77] + 0.05, steps=2, lr=0.6))
    l2s = dev.lambda_2_sqrt
    assert all(np.all(np.diagonal(l2s[l]) < 0) for l in range(L))
    dev.close()


def test_sites_round_trip_and_gradients_for_two_latents():
    import tsvgp_b200 as tb
    rng = np.random.default_rng(3)
    N, M, D, L = 600, 40, 3, 2
    X = rng.standard_normal((N, D))
    Z = X[:M].copy()
    Y = np.stack([np.sin(X.sum(1)), np.cos(X[:, 0])], axis=1) + 0.1 * rng.standard_normal((N, L))
    kernel, lik = orc.SquaredExponential(variance=0.9, lengthscales=0.8), orc.Gaussian(variance=0.2)    # cond(Kuu) ~ 1e3
    ref = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(Z.copy()), num_latent_gps=L)
    ref.natgrad_step((X, Y), lr=0.7)
    dev = tb.t_SVGP(kernel, lik, Z.copy(), lambda_1=ref.lambda_1, lambda_2_sqrt=ref.lambda_2_sqrt)   # [M, L] / [L, M, M] in
    assert dev.num_latent_gps == L
    np.testing.assert_array_equal(dev.lambda_1, ref.lambda_1)
    np.testing.assert_array_equal(dev.lambda_2_sqrt, ref.lambda_2_sqrt)
    # M-step gradients add up over the latents (shared kernel)
    e_dev, g_dev = dev.elbo_and_grad((X, Y))
    tot = {"variance": 0.0, "lengthscales": 0.0, "Z": 0.0, "likelihood": 0.0}
    mag = dict.fromkeys(tot, 0.0)      # the latents' contributions may cancel: errors are read against their summed magnitudes
    for l in range(L):
        one = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(Z.copy()), lambda_1=ref.lambda_1[:, l:l + 1], lambda_2_sqrt=ref.lambda_2_sqrt[l:l + 1])
        _, g = orc.elbo_gradients(one, (X, Y[:, l:l + 1]))
        for k in tot:
            tot[k] = tot[k] + g[k]
            mag[k] += float(np.max(np.abs(g[k])))
    errs = {k: float(np.max(np.abs(np.asarray(g_dev[k]) - tot[k])) / mag[k]) for k in tot}
    errs["elbo"] = abs(e_dev - ref.elbo((X, Y))) / abs(ref.elbo((X, Y)))
    check(errs)
    dev.close()


def test_softmax_monte_carlo_with_fixed_draws_matches_the_oracle():
    import tsvgp_b200 as tb
    rng = np.random.default_rng(5)
    N, M, D, C, S = 900, 64, 4, 5, 40
    X = rng.standard_normal((N, D))
    Z = X[:M].copy()
    W = rng.standard_normal((D, C))
    labels = np.argmax(X @ W + 0.5 * rng.standard_normal((N, C)), axis=1).astype(float)[:, None]
    eps = rng.standard_normal((S, N, C))
    kernel = orc.Matern52(variance=1.0, lengthscales=np.full(D, 1.7))      # ARD, as mnist.py:117 uses
    lik_ref = orc.Softmax(C, num_monte_carlo_points=S, epsilon=eps)
    ref = orc.OracleTSVGP(kernel, lik_ref, orc.InducingPoints(Z.copy()), num_latent_gps=C, num_data=3 * N)
    dev = tb.t_SVGP(kernel, lik_ref, Z.copy(), num_latent_gps=C, num_data=3 * N)
    dev.set_option("chunk", 256)
    dev.set_mc_epsilon(eps)
    errs = {}
    for s in range(2):
        e_ref = ref.elbo((X, labels))
        e_dev = dev.natgrad_step((X, labels), lr=0.4, return_elbo=True)
        ref.natgrad_step((X, labels), lr=0.4)
        errs[f"elbo_before_step{s}"] = abs(e_dev - e_ref) / abs(e_ref)
        errs[f"lambda_1_step{s}"] = relerr(dev.lambda_1, ref.lambda_1)
        errs[f"lambda_2_step{s}"] = relerr(dev.lambda_2, ref.lambda_2)
    mu_d, var_d = dev.predict_f(X[:50])
    mu_r, var_r = ref.predict_f(X[:50])
    errs["mean"], errs["var"] = relerr(mu_d, mu_r), relerr(var_d, var_r)
    check(errs)
    # the model learns: predicted class = argmax of the latent means on most training points
    acc = np.mean(np.argmax(dev.predict_f(X)[0], axis=1) == labels[:, 0])
    assert acc > 0.6, acc
    # the built-in generator (no explicit draws): same objective up to Monte-Carlo error, a new draw per call
    dev.set_mc_epsilon(None)
    e1, e2 = dev.elbo((X, labels)), dev.elbo((X, labels))
    e_fixed = ref.elbo((X, labels))
    assert e1 != e2 and abs(e1 - e_fixed) < 0.05 * abs(e_fixed) and abs(e2 - e_fixed) < 0.05 * abs(e_fixed)
    dev.close()


def test_shared_independent_inducing_variables_are_the_shared_Z():
    # reference tsvgp.py:249-254 ("hack to get heterokedastic demo to run"): SharedIndependentInducingVariables wraps ONE InducingPoints
    # object shared by all latents; on this path that is exactly num_latent_gps = L over a shared Z
    import tsvgp_b200 as tb
    from tsvgp_b200 import standins as st
    rng = np.random.default_rng(2)
    N, M, D, L = 500, 48, 2, 2
    X = rng.standard_normal((N, D))
    Z = X[:M].copy()
    Y = np.stack([np.sin(X[:, 0]), np.cos(X[:, 1])], 1) + 0.1 * rng.standard_normal((N, L))
    kernel, lik = orc.SquaredExponential(variance=1.0, lengthscales=0.9), orc.Gaussian(variance=0.1)
    a = tb.t_SVGP(kernel, lik, Z.copy(), num_latent_gps=L)
    b = tb.t_SVGP(kernel, lik, st.SharedIndependentInducingVariables(st.InducingPoints(Z.copy())), num_latent_gps=L)
    for m in (a, b):
        m.natgrad_step((X, Y), lr=0.8)
    np.testing.assert_array_equal(a.lambda_1, b.lambda_1)
    np.testing.assert_array_equal(a.lambda_2_sqrt, b.lambda_2_sqrt)
    a.close(); b.close()
