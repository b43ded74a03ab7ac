"""
GPU: the whitened sibling `t_SVGP_white` (second "next" row of SURVEY 8f) through the C-ABI against the oracle's restatement
(oracle.OracleTSVGPWhite, pinned by the reference's property tests in tests/test_oracle_white.py).  Tolerance 1e-9 norm-wise.
"""
import numpy as np
import pytest

from oracle import tsvgp_oracle as orc

pytestmark = pytest.mark.gpu


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("name,n,M,num_data", [("cfg1", 3000, 50, None), ("cfg2", 2500, 200, 25_000), ("cfg3", 3000, 384, 30_000),
                                               ("cfg2", 1000, 129, None)])
def test_white_model_matches_oracle(name, n, M, num_data):
    import tsvgp_b200 as tb
    import tsvgp_b200.synth as synth
    cfg = synth.describe(name)
    X, Y, Z = synth.make_minibatch(cfg, n_rows=n, M=M)
    kernel, lik = synth.build_objects(cfg, orc)
    ref = orc.OracleTSVGPWhite(kernel, lik, orc.InducingPoints(Z.copy()), num_data=num_data)
    dev = tb.t_SVGP_white(kernel, lik, Z.copy(), num_data=num_data)
    errs = {}
    for s in range(2):
        e_ref = ref.elbo((X, Y))
        e_dev = dev.natgrad_step((X, Y), lr=cfg["lr"], return_elbo=True)
        ref.natgrad_step((X, Y), lr=cfg["lr"])
        errs[f"elbo_before{s}"] = abs(e_dev - e_ref) / abs(e_ref)
        errs[f"lambda_1_{s}"] = relerr(dev.lambda_1, ref.lambda_1)
        errs[f"lambda_2_{s}"] = relerr(dev.lambda_2, ref.lambda_2)
    mu_d, var_d = dev.predict_f(X[:200] + 0.05)
    mu_r, var_r = ref.predict_f(X[:200] + 0.05)
    errs["mean"], errs["var"] = relerr(mu_d, mu_r), relerr(var_d, var_r)
    errs["prior_kl"] = abs(dev.prior_kl() - ref.prior_kl()) / abs(ref.prior_kl())
    m_d, cs_d = dev.get_mean_chol_cov_inducing_posterior()
    m_r, cs_r = ref.get_mean_chol_cov_inducing_posterior()
    errs["m_q"], errs["S_q"] = relerr(m_d, m_r), relerr(cs_d[0] @ cs_d[0].T, cs_r[0] @ cs_r[0].T)
    bad = {k: v for k, v in errs.items() if not v <= 1e-9}
    assert not bad, (bad, errs)
    with pytest.raises(AttributeError):
        dev.lambda_2_sqrt
    dev.close()


def test_white_and_plain_models_agree():
    # reference tests/models/test_tsvgp_white.py:58-85 (decimal=4): same ELBO at init, same predictions after one step
    import tsvgp_b200 as tb
    rng = np.random.RandomState(123)
    X = rng.rand(8, 1) * 2 - 1
    Y = np.sin(X * 3 * 3.14) + 0.3 * np.cos(X * 9 * 3.14) + 0.5 * np.sin(X * 7 * 3.14) + 0.2 * rng.randn(8, 1)
    kernel, lik = orc.SquaredExponential(lengthscales=2.0, variance=2.25), orc.Gaussian(variance=0.3)
    a, b = tb.t_SVGP(kernel, lik, X.copy()), tb.t_SVGP_white(kernel, lik, X.copy())
    np.testing.assert_almost_equal(a.elbo((X, Y)), b.elbo((X, Y)), decimal=4)
    a.natgrad_step((X, Y), lr=0.9); b.natgrad_step((X, Y), lr=0.9)
    (ma, va), (mb, vb) = a.predict_f(X), b.predict_f(X)
    np.testing.assert_array_almost_equal(ma, mb, decimal=4)
    np.testing.assert_array_almost_equal(va, vb, decimal=4)
    a.close(); b.close()


def test_white_model_fails_like_the_reference_when_lambda_2_turns_indefinite():
    # Student-t: the variance gradient is not clipped in tsvgp_white.py:183-212, Lambda_2 can lose positive definiteness and the
    # reference's next Cholesky raises; the device reports the same condition as NotPositiveDefiniteError at the same step
    import numpy.linalg
    import tsvgp_b200 as tb
    import tsvgp_b200.synth as synth
    cfg = synth.describe("cfg5")
    X, Y, Z = synth.make_minibatch(cfg, n_rows=2000, M=256)
    kernel, lik = synth.build_objects(cfg, orc)
    ref = orc.OracleTSVGPWhite(kernel, lik, orc.InducingPoints(Z.copy()), num_data=50_000)
    dev = tb.t_SVGP_white(kernel, lik, Z.copy(), num_data=50_000)
    failed_ref = failed_dev = None
    for step in range(4):
        try:
            ref.natgrad_step((X, Y), lr=cfg["lr"])
        except (numpy.linalg.LinAlgError, FloatingPointError):
            failed_ref = step
        try:
            dev.natgrad_step((X, Y), lr=cfg["lr"])
        except tb.InvalidArgumentError:
            failed_dev = step
        if failed_ref is not None or failed_dev is not None:
            break
        assert relerr(dev.lambda_1, ref.lambda_1) < 1e-9 and relerr(dev.lambda_2, ref.lambda_2) < 1e-9
    assert failed_ref == failed_dev
    dev.close()


@pytest.mark.parametrize("jitter", [1e-6, 0.0])
def test_predict_f_extra_data(jitter):
    # tsvgp_white.py:134-158 (the reference's test_condit.py:101 calls it with jitter = 0): conditioning on extra data for the
    # prediction only; the sites are left as they were
    import tsvgp_b200 as tb
    import tsvgp_b200.synth as synth
    cfg = synth.describe("cfg3")
    X, Y, Z = synth.make_minibatch(cfg, n_rows=3000, M=160)
    kernel, lik = synth.build_objects(cfg, orc)
    ref = orc.OracleTSVGPWhite(kernel, lik, orc.InducingPoints(Z.copy()))
    dev = tb.t_SVGP_white(kernel, lik, Z.copy())
    ref.natgrad_step((X[:1500], Y[:1500]), lr=0.6)
    dev.natgrad_step((X[:1500], Y[:1500]), lr=0.6)
    l1, l2 = dev.lambda_1, dev.lambda_2
    mu_r, var_r = ref.predict_f_extra_data(X[:300] + 0.1, (X[1500:], Y[1500:]), jitter=jitter)
    mu_d, var_d = dev.predict_f_extra_data(X[:300] + 0.1, (X[1500:], Y[1500:]), jitter=jitter)
    assert relerr(mu_d, mu_r) < 1e-9 and relerr(var_d, var_r) < 1e-9
    np.testing.assert_array_equal(dev.lambda_1, l1)
    np.testing.assert_array_equal(dev.lambda_2, l2)
    mu0_d, var0_d = dev.predict_f(X[:300] + 0.1)                # and ordinary predictions are back to the unconditioned ones
    mu0_r, var0_r = ref.predict_f(X[:300] + 0.1)
    assert relerr(mu0_d, mu0_r) < 1e-9 and relerr(var0_d, var0_r) < 1e-9
    dev.close()


@pytest.mark.parametrize("lik_name,L,M", [("gaussian", 3, 200), ("bernoulli", 2, 130)])
def test_white_model_with_several_latents(lik_name, L, M):
    # tsvgp_white.py:79-89 builds one Lambda_2 per latent, util.py:60-88 / :264-291 / :411-426 loop over them and the update
    # (:240-246) broadcasts K_uu over the latent axis: L sets of sites in ONE context, one Kuf slab and one |LA^-1 k|^2 pass for all
    import tsvgp_b200 as tb
    import tsvgp_b200.synth as synth
    rng = np.random.default_rng(11)
    n = 1500
    cfg = synth.describe("cfg3" if lik_name == "gaussian" else "cfg2")      # inputs with cond(Kuu) ~ 1e3, as in the L = 1 cases above
    X, _, Z = synth.make_minibatch(cfg, n_rows=n, M=M)
    kernel, lik = synth.build_objects(cfg, orc)
    F = np.stack([np.sin(X @ rng.standard_normal(X.shape[1])) for _ in range(L)], 1)
    Y = F + 0.3 * rng.standard_normal((n, L)) if lik_name == "gaussian" else (F + 0.3 * rng.standard_normal((n, L)) > 0).astype(float)
    ref = orc.OracleTSVGPWhite(kernel, lik, orc.InducingPoints(Z.copy()), num_latent_gps=L, num_data=4 * n)
    dev = tb.t_SVGP_white(kernel, lik, Z.copy(), num_latent_gps=L, num_data=4 * n)
    dev.set_option("chunk", 512)                     # several slabs per stream
    errs = {}
    for s in range(2):
        e_ref = ref.elbo((X, Y))
        e_dev = dev.natgrad_step((X, Y), lr=0.5, return_elbo=True)
        ref.natgrad_step((X, Y), lr=0.5)
        errs[f"elbo_before{s}"] = abs(e_dev - e_ref) / abs(e_ref)
        errs[f"lambda_1_{s}"] = relerr(dev.lambda_1, ref.lambda_1)
        errs[f"lambda_2_{s}"] = relerr(dev.lambda_2, ref.lambda_2)
    assert dev.lambda_1.shape == (M, L) and dev.lambda_2.shape == (L, M, M)
    mu_d, var_d = dev.predict_f(X[:200] + 0.05)
    mu_r, var_r = ref.predict_f(X[:200] + 0.05)
    assert mu_d.shape == (200, L)
    errs["mean"], errs["var"] = relerr(mu_d, mu_r), relerr(var_d, var_r)
    errs["prior_kl"] = abs(dev.prior_kl() - ref.prior_kl()) / abs(ref.prior_kl())
    m_d, cs_d = dev.get_mean_chol_cov_inducing_posterior()
    m_r, cs_r = ref.get_mean_chol_cov_inducing_posterior()
    errs["m_q"] = relerr(m_d, m_r)
    errs["S_q"] = max(relerr(cs_d[l] @ cs_d[l].T, cs_r[l] @ cs_r[l].T) for l in range(L))
    # conditioning on extra data, latent by latent, leaves every latent's sites as they were
    l1, l2 = dev.lambda_1, dev.lambda_2
    Xe, Ye = X[:600] * 0.9 + 0.1, Y[:600]
    mu_r, var_r = ref.predict_f_extra_data(X[:100] + 0.1, (Xe, Ye))
    mu_d, var_d = dev.predict_f_extra_data(X[:100] + 0.1, (Xe, Ye))
    errs["extra_mean"], errs["extra_var"] = relerr(mu_d, mu_r), relerr(var_d, var_r)
    np.testing.assert_array_equal(dev.lambda_1, l1)
    np.testing.assert_array_equal(dev.lambda_2, l2)
    # assigned sites round-trip in the reference's layouts
    ref2 = orc.OracleTSVGPWhite(kernel, lik, orc.InducingPoints(Z.copy()), lambda_1=l1, lambda_2=l2)
    dev2 = tb.t_SVGP_white(kernel, lik, Z.copy(), lambda_1=l1, lambda_2=l2)
    assert dev2.num_latent_gps == L
    errs["elbo_assigned"] = abs(dev2.elbo((X, Y)) - ref2.elbo((X, Y))) / abs(ref2.elbo((X, Y)))
    bad = {k: v for k, v in errs.items() if not v <= 1e-9}
    assert not bad, (bad, errs)
    dev.close(); dev2.close()
