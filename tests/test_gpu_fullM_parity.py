"""
GPU oracle parity at the inducing-point counts that are BENCHMARKED (VERDICT r01 "what's weak" #2): the reduced-M cases of
tests/test_gpu_parity.py never run the 16/32/64-block Cholesky recursion, the split-K thresholds at large M, the two-piece balanced
SYRK or several slabs per stream.  Here:
  cfg3  M = 2048 (full), 40 960 rows = 4 slabs of 10 240 over 2 streams -> reference-order oracle, tolerance 1e-9
  cfg5  M = 4096 (full), 8 192 rows, Student-t GH-20                    -> reference-order oracle, tolerance 1e-9
  cfg4  M = 8192 (full), 16 384 rows -> tests/algo_model.py (8 M^3 instead of the oracle's 37 M^3; pinned to the oracle at
        M <= 1024 by tests/test_algebra_model.py on the CPU and by test_algo_model_is_the_oracle_at_m1024 below), tolerance 1e-9
Reference being matched: src/models/tsvgp.py:234-304 (natgrad_step), :79-95 (elbo), :97-114 (predict_f).
"""
import numpy as np
import pytest

from oracle import tsvgp_oracle as orc
from tests import algo_model as am
from tests.test_gpu_parity import check, relerr, run_pair

pytestmark = pytest.mark.gpu


def test_cfg3_full_m2048_several_slabs_per_stream():
    import tsvgp_b200.synth as synth
    cfg = synth.describe("cfg3")
    check(run_pair(cfg, n_rows=40_960, M=2048, steps=2, num_data=409_600, Xtest_rows=1024))


def test_cfg5_full_m4096_student_t():
    import tsvgp_b200.synth as synth
    cfg = synth.describe("cfg5")
    check(run_pair(cfg, n_rows=8_192, M=4096, steps=2, num_data=204_800, Xtest_rows=512))


def _algo_step(kernel, noise_var, X, Y, Z, l1, L2, lr, scale):
    """One natgrad_step in the device's algebra (tests/algo_model.py), Gaussian likelihood; returns the new sites, the ELBO
    before the step and the predictive moments at X under the OLD sites."""
    K = kernel.K(Z)
    Kuf = kernel.K(Z, X)
    pre = am.prepare(K, l1, L2)
    mu, var = am.marginals(Kuf, kernel.K_diag(X), pre["T"], pre["alpha"])
    y = Y[:, 0]
    mu = mu[:, 0]
    ve = -0.5 * np.log(2 * np.pi) - 0.5 * np.log(noise_var) - 0.5 * ((y - mu) ** 2 + var) / noise_var
    g = (y - mu) / noise_var
    h = np.full_like(g, min(-0.5 / noise_var, -1e-8))
    elbo = scale * np.sum(ve) - am.kl(pre["K6"], pre["T"], pre["alpha"], pre["Uw"])
    n1, nL2 = am.natgrad(K, Kuf, g, h, pre["alpha"], l1, L2, lr, scale)
    return n1, nL2, elbo, mu, var


def test_algo_model_is_the_oracle_at_m1024():
    # the arbiter of the M = 8192 test below, pinned to the reference-order oracle at the largest M the oracle handles quickly
    import tsvgp_b200.synth as synth
    cfg = synth.describe("cfg4")
    X, Y, Z = synth.make_minibatch(cfg, n_rows=3000, M=1024)
    kernel, lik = synth.build_objects(cfg, orc)
    ref = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(Z.copy()))
    ref.natgrad_step((X, Y), lr=0.5)
    l1, L2 = ref.lambda_1.copy(), ref.lambda_2_sqrt[0].copy()
    e_ref = ref.elbo((X, Y))
    n1, nL2, elbo, mu, var = _algo_step(kernel, 0.1, X, Y, Z, l1, L2, 0.5, 1.0)
    mu_r, var_r = ref.predict_f(X)
    ref.natgrad_step((X, Y), lr=0.5)
    errs = {"lambda_1": relerr(n1, ref.lambda_1), "lambda_2": relerr(nL2 @ nL2.T, ref.lambda_2[0]), "elbo": abs(elbo - e_ref) / abs(e_ref),
            "mean": relerr(mu, mu_r[:, 0]), "var": relerr(var, var_r[:, 0])}
    check(errs)


def test_cfg4_full_m8192_against_the_algebra_model():
    import tsvgp_b200 as tb
    import tsvgp_b200.synth as synth
    cfg = synth.describe("cfg4")
    n_rows = 16_384
    X, Y, Z = synth.make_minibatch(cfg, n_rows=n_rows, M=8192)
    kernel, lik = synth.build_objects(cfg, orc)
    dev = tb.t_SVGP(kernel, lik, Z.copy())
    M = Z.shape[0]
    l1, L2 = np.zeros((M, 1)), -1e-10 * np.eye(M)      # tsvgp.py:174-180
    errs = {}
    for s in range(2):
        n1, nL2, elbo, mu, var = _algo_step(kernel, 0.1, X, Y, Z, l1, L2, cfg["lr"], 1.0)
        if s == 1:                                     # predictive moments under non-trivial sites
            mu_d, var_d = dev.predict_f(X[:2048])
            errs["mean"], errs["var"] = relerr(mu_d[:, 0], mu[:2048]), relerr(var_d[:, 0], var[:2048])
        e_dev = dev.natgrad_step((X, Y), lr=cfg["lr"], return_elbo=True)
        errs[f"elbo_before_step{s}"] = abs(e_dev - elbo) / abs(elbo)
        errs[f"lambda_1_step{s}"] = relerr(dev.lambda_1, n1)
        errs[f"lambda_2_step{s}"] = relerr(dev.lambda_2[0], nL2 @ nL2.T)
        l1, L2 = n1, nL2
    check(errs)
    dev.close()
