"""
GPU, BASELINE.json's full sizes (where the oracle cannot run in seconds): size-independent properties of the path.
  * Gaussian likelihood, lr = 1: one natgrad step reaches the optimum for the minibatch, so a second step is a fixed point.
  * additivity: the step on a minibatch equals the step on the same rows presented in a different slab schedule.
  * the ELBO by-product of natgrad_step equals a separate elbo() call on the same state.
  * predict_f at the inducing inputs reproduces K6^-1-weighted posterior mean m_q up to the two-jitter quirk (1e-6).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


def test_cfg3_full_minibatch_properties():
    import tsvgp_b200 as tb
    import tsvgp_b200.synth as synth
    from tsvgp_b200 import standins as st
    cfg = synth.describe("cfg3")                       # M = 2048, D = 16, Matern-5/2, minibatch 1e6
    X, Y, Z = synth.make_minibatch(cfg, n_rows=cfg["Nb"], M=cfg["M"])
    kernel, lik = synth.build_objects(cfg, st)
    m = tb.t_SVGP(kernel, lik, Z, num_data=cfg["N"])
    Xd, Yd = m.device_array(X), m.device_array(Y)
    m.set_data((Xd, Yd))
    e0 = m.natgrad_step(lr=1.0, return_elbo=True)
    assert np.isfinite(e0)
    l1, l2 = m.lambda_1, m.lambda_2
    e1 = m.natgrad_step(lr=1.0, return_elbo=True)      # fixed point of the Gaussian update (reference test_tsvgp.py:134-145)
    assert relerr(m.lambda_1, l1) < 1e-7 and relerr(m.lambda_2, l2) < 1e-7
    assert abs(m.elbo() - e1) < 1e-9 * abs(e1)         # by-product ELBO == elbo() on the same (fixed-point) state
    assert e1 > e0
    # a different slab schedule (chunk 4096, one stream) is the same step
    m2 = tb.t_SVGP(kernel, lik, Z, num_data=cfg["N"])
    m2.set_option("chunk", 4096); m2.set_option("streams", 1)
    m2.set_data((Xd, Yd))
    m2.natgrad_step(lr=1.0)
    assert relerr(m2.lambda_1, l1) < 1e-10 and relerr(m2.lambda_2, l2) < 1e-10
    mu, var = m.predict_f(X[:4096])
    assert np.all(var > 0) and np.all(var < 1.0 + 1e-12)          # posterior variance within the prior variance
    assert np.sqrt(np.mean((mu - Y[:4096]) ** 2)) < 0.5           # fits the data to about the noise level (sigma = 0.316)
    m.close(); m2.close()


def test_cfg5_quadrature_step_improves_elbo_at_scale():
    import tsvgp_b200 as tb
    import tsvgp_b200.synth as synth
    from tsvgp_b200 import standins as st
    cfg = synth.describe("cfg5")                       # M = 4096, D = 32, Student-t GH-20; 200k of the 2M-row minibatch
    X, Y, Z = synth.make_minibatch(cfg, n_rows=200_000, M=cfg["M"])
    kernel, lik = synth.build_objects(cfg, st)
    m = tb.t_SVGP(kernel, lik, Z, num_data=cfg["N"])
    m.set_data((m.device_array(X), m.device_array(Y)))
    elbos = [m.natgrad_step(lr=cfg["lr"], return_elbo=True) for _ in range(4)]
    elbos.append(m.elbo())
    assert all(b > a for a, b in zip(elbos, elbos[1:])), elbos    # damped natural-gradient ascent on a fixed minibatch
    l2s = m.lambda_2_sqrt[0]
    assert np.all(np.diagonal(l2s) < 0) and np.all(np.isfinite(l2s))
    m.close()


def test_cfg4_m8192_dense_phase_properties():
    # the Cholesky-bound config at its full M = 8192 (64 tile rows: deepest blocked Cholesky / triangular-inverse recursion the
    # library runs) on a reduced minibatch: the Gaussian lr = 1 fixed point, ELBO consistency, finite negative-diagonal factor
    import tsvgp_b200 as tb
    import tsvgp_b200.synth as synth
    from tsvgp_b200 import standins as st
    cfg = synth.describe("cfg4")
    X, Y, Z = synth.make_minibatch(cfg, n_rows=16_384, M=cfg["M"])
    kernel, lik = synth.build_objects(cfg, st)
    m = tb.t_SVGP(kernel, lik, Z, num_data=cfg["N"])
    m.set_data((m.device_array(X), m.device_array(Y)))
    e0 = m.natgrad_step(lr=1.0, return_elbo=True)
    l1 = m.lambda_1
    e1 = m.natgrad_step(lr=1.0, return_elbo=True)
    assert e1 > e0
    assert relerr(m.lambda_1, l1) < 1e-6                          # one full Gaussian step is already the optimum
    assert abs(m.elbo() - e1) < 1e-8 * abs(e1)
    d = np.diagonal(m.lambda_2_sqrt[0])
    assert np.all(np.isfinite(d)) and np.all(d < 0)
    mu, var = m.predict_f(X[:2048])
    assert np.all(var > 0) and np.sqrt(np.mean((mu - Y[:2048]) ** 2)) < 0.6
    m.close()
