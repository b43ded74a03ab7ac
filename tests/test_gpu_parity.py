"""
GPU parity: the CUDA path (through the C-ABI, via the Python mirror of t_SVGP) against the NumPy oracle on the same
seeded inputs.  Tolerance (BASELINE.json north_star): relative error <= 1e-9 on lambda_1, lambda_2, ELBO and the
predictive moments — norm-wise, max|a-b| / max|b| (element-wise relative error on near-zero lambda_2 entries
is meaningless).
"""
import numpy as np
import pytest

from oracle import tsvgp_oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-9


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def _objects(cfg):
    import tsvgp_b200.synth as synth
    return synth.build_objects(cfg, orc)


def run_pair(cfg, n_rows, M, steps=2, num_data=None, lr=None, Xtest_rows=257, mean_function=None, ard=None, options=None):
    import tsvgp_b200 as tb
    import tsvgp_b200.synth as synth

    X, Y, Z = synth.make_minibatch(cfg, n_rows=n_rows, M=M)
    kernel, lik = _objects(cfg)
    if ard is not None:
        kernel.lengthscales = orc._param(ard)
    lr = cfg["lr"] if lr is None else lr
    ref = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(Z.copy()), num_data=num_data, mean_function=mean_function)
    dev = tb.t_SVGP(kernel, lik, orc.InducingPoints(Z.copy()), num_data=num_data, mean_function=mean_function)
    for k, v in (options or {}).items():
        dev.set_option(k, v)
    rng = np.random.default_rng(7)
    Xt = X[rng.permutation(X.shape[0])[:Xtest_rows]] + 0.05 * rng.standard_normal((min(Xtest_rows, X.shape[0]), X.shape[1]))
    errs = {}
    for s in range(steps):
        e_ref = ref.elbo((X, Y))
        e_dev_by = dev.natgrad_step((X, Y), lr=lr, return_elbo=True)
        ref.natgrad_step((X, Y), lr=lr)
        errs[f"elbo_before_step{s}"] = abs(e_dev_by - e_ref) / abs(e_ref)
        errs[f"lambda_1_step{s}"] = relerr(dev.lambda_1, ref.lambda_1)
        errs[f"lambda_2_step{s}"] = relerr(dev.lambda_2, ref.lambda_2)
    errs["elbo"] = abs(dev.elbo((X, Y)) - ref.elbo((X, Y))) / abs(ref.elbo((X, Y)))
    mu_d, var_d = dev.predict_f(Xt)
    mu_r, var_r = ref.predict_f(Xt)
    errs["mean"] = relerr(mu_d, mu_r)
    errs["var"] = relerr(var_d, var_r)
    errs["prior_kl"] = abs(dev.prior_kl() - ref.prior_kl()) / max(abs(ref.prior_kl()), 1e-300)
    m_d, cs_d = dev.get_mean_chol_cov_inducing_posterior()
    m_r, cs_r = ref.get_mean_chol_cov_inducing_posterior()
    errs["m_q"] = relerr(m_d, m_r)
    errs["S_q"] = relerr(cs_d[0] @ cs_d[0].T, cs_r[0] @ cs_r[0].T)
    errs["_route"] = 0.0 * dev.timings()["route"]
    l2s = dev.lambda_2_sqrt
    assert np.all(np.diagonal(l2s[0]) < 0) and np.allclose(np.triu(l2s[0], 1), 0.0)  # tsvgp.py:300 : -chol(...)
    dev.close()
    return errs


def _record(errs):
    import inspect, json, os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        caller = inspect.stack()[2].function
        with open(os.path.join(out, "parity_errors.jsonl"), "a") as f:
            f.write(json.dumps({"test": caller, "max": max(errs.values()), "errs": errs}) + "\n")


def check(errs, tol=TOL):
    _record(errs)
    bad = {k: v for k, v in errs.items() if not (v <= tol)}
    assert not bad, "relative errors above %.0e: %s\nall: %s" % (tol, bad, errs)


def test_cfg1_gaussian_se_1d_full():
    import tsvgp_b200.synth as synth
    cfg = synth.describe("cfg1")
    check(run_pair(cfg, n_rows=10_000, M=50, steps=2))   # the reference's own CPU-runnable case at its full size


def test_cfg2_bernoulli_gh20():
    import tsvgp_b200.synth as synth
    cfg = synth.describe("cfg2")
    check(run_pair(cfg, n_rows=10_000, M=500, steps=2, num_data=100_000))   # full minibatch size of configs[1]


def test_cfg3_matern52_reduced():
    import tsvgp_b200.synth as synth
    cfg = synth.describe("cfg3")
    check(run_pair(cfg, n_rows=6000, M=640, steps=2, num_data=60_000))


def test_cfg4_se_reduced():
    import tsvgp_b200.synth as synth
    cfg = synth.describe("cfg4")
    check(run_pair(cfg, n_rows=5000, M=768, steps=2))


def test_cfg5_student_t_reduced():
    import tsvgp_b200.synth as synth
    cfg = synth.describe("cfg5")
    check(run_pair(cfg, n_rows=5000, M=512, steps=2, num_data=125_000))


@pytest.mark.parametrize("n_rows,M", [(1, 1), (3, 2), (127, 128), (129, 129), (1000, 257), (2049, 100)])
def test_ragged_shapes(n_rows, M):
    # rows / inducing counts that are not multiples of the 128-wide tiles, down to a single point
    import tsvgp_b200.synth as synth
    cfg = synth.describe("cfg2", ls=2.0)
    check(run_pair(cfg, n_rows=n_rows, M=M, steps=2, Xtest_rows=min(n_rows, 50)))


def test_multi_slab_and_single_stream():
    # several slabs per pass (ping-pong accumulators) and the single-stream schedule give the same numbers
    import tsvgp_b200.synth as synth
    cfg = synth.describe("cfg3")
    check(run_pair(cfg, n_rows=3000, M=256, steps=2, options={"chunk": 256}))
    check(run_pair(cfg, n_rows=3000, M=256, steps=2, options={"chunk": 384, "streams": 1}))


def test_whitened_route_matches_too():
    # the reference-order (whitened) statistics route, forced, on well-conditioned inputs
    import tsvgp_b200.synth as synth
    check(run_pair(synth.describe("cfg3"), n_rows=3000, M=384, steps=2, num_data=30_000, options={"route": 2}))
    check(run_pair(synth.describe("cfg2"), n_rows=2000, M=200, steps=2, options={"route": 2, "chunk": 512}))


def test_ill_conditioned_inducing_points_take_the_whitened_route():
    # cond(Kuu) ~ 1e7 (long lengthscale): the automatic route must whiten; parity is stated at the reference's own
    # rounding level for this conditioning (eps * cond ~ 1e-9 .. 1e-8), not gated at 1e-9  
    import tsvgp_b200 as tb
    import tsvgp_b200.synth as synth
    cfg = synth.describe("cfg2", ls=4.0)
    errs = run_pair(cfg, n_rows=2000, M=300, steps=2)
    check(errs, tol=1e-6)


def test_exact_route_matches_too():
    # route 3: A = K9^-1 Kuf formed per slab and the Gram product A^T diag(h) A accumulated directly (tsvgp.py:271-281 literally)
    import tsvgp_b200.synth as synth
    check(run_pair(synth.describe("cfg3"), n_rows=3000, M=384, steps=2, num_data=30_000, options={"route": 3}))
    check(run_pair(synth.describe("cfg2"), n_rows=2000, M=200, steps=2, options={"route": 3, "chunk": 512}))


def test_extreme_conditioning_keeps_the_reference_error_behaviour():
    # examples/c_api_example.c's data: 40 inducing points 0.05 apart under an SE kernel of lengthscale 0.2 — cond(Kuu) = 5e17,
    # cond(Kuu + 1e-9 I) = 9e9.  The reference's order (A first, then the Gram product) keeps -2 Lambda_2 + jitter I positive
    # definite here and the step succeeds; the two-sided product of the whitened route does not (min eigenvalue -7e-4).  The
    # automatic route must therefore take the exact route and must not raise.  Results agree with the oracle at the level two
    # float64 implementations can agree at this conditioning (the oracle itself is ~1e-3 from exact arithmetic here).
    import tsvgp_b200 as tb
    N, M = 2000, 40
    s = 12345
    X, Y = np.zeros((N, 1)), np.zeros((N, 1))
    for i in range(N):
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        X[i] = 2.0 * (s >> 8) / 16777216.0 - 1.0
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        Y[i] = np.sin(6.0 * X[i]) + 0.2 * ((s >> 8) / 16777216.0 - 0.5)
    Z = np.linspace(-1.0, 1.0, M)[:, None]
    k, lik = orc.SquaredExponential(lengthscales=0.2, variance=1.0), orc.Gaussian(variance=0.05)
    dev = tb.t_SVGP(k, lik, orc.InducingPoints(Z.copy()))
    ref = orc.OracleTSVGP(k, lik, orc.InducingPoints(Z.copy()))
    for _ in range(3):
        dev.natgrad_step((X, Y), lr=0.9)
        ref.natgrad_step((X, Y), lr=0.9)
    assert dev.timings()["route"] == 3 and dev.timings()["cond_est"] > 1e8
    Xs = np.linspace(-0.9, 0.9, 50)[:, None]
    mu_d, var_d = dev.predict_f(Xs)
    mu_r, var_r = ref.predict_f(Xs)
    assert relerr(mu_d, mu_r) < 1e-3 and np.max(np.abs(mu_d[:, 0] - np.sin(6.0 * Xs[:, 0]))) < 0.05
    dev.close()


def test_lr_one_and_ard_and_mean_function():
    import tsvgp_b200.synth as synth
    cfg = synth.describe("cfg5")
    ard = np.linspace(2.0, 4.0, cfg["D"])
    check(run_pair(cfg, n_rows=1500, M=200, steps=3, lr=1.0, ard=ard, mean_function=lambda X: 0.3 + 0.1 * X[:, :1]))


def test_minibatch_scaling_identity():
    # reference tests/models/test_tsvgp.py:148-165 : num_data = 8 with one point == the point replicated 8 times
    import tsvgp_b200 as tb
    rng = np.random.RandomState(123)
    X = rng.rand(8, 1) * 2 - 1
    Y = np.sin(3 * X) + 0.2 * rng.randn(8, 1)
    k, lik = orc.SquaredExponential(lengthscales=0.5, variance=2.25), orc.Gaussian(variance=0.3)
    m = tb.t_SVGP(k, lik, orc.InducingPoints(X.copy()))
    for _ in range(5):
        m.natgrad_step((X, Y), lr=0.9)
    e2 = m.elbo((X[0].repeat(8)[:, None], Y[0].repeat(8)[:, None]))
    m.num_data = 8
    e1 = m.elbo((X[0][:, None], Y[0][:, None]))
    assert abs(e1 - e2) <= 1e-10 * abs(e2)
    m.close()


def test_gaussian_fixed_point_is_gp_regression():
    # reference tests/models/test_tsvgp.py:106-120 (decimal=4): Z = X, lr -> optimum == exact GP regression
    import tsvgp_b200 as tb
    rng = np.random.RandomState(123)
    X = rng.rand(8, 1) * 2 - 1
    Y = np.sin(X * 3 * 3.14) + 0.3 * np.cos(X * 9 * 3.14) + 0.5 * np.sin(X * 7 * 3.14) + 0.2 * rng.randn(8, 1)
    k, lik = orc.SquaredExponential(lengthscales=2.0, variance=2.25), orc.Gaussian(variance=0.3)
    m = tb.t_SVGP(k, lik, orc.InducingPoints(X.copy()))
    for _ in range(10):
        m.natgrad_step((X, Y), lr=0.9)
    np.testing.assert_almost_equal(m.elbo((X, Y)), orc.gpr_log_marginal_likelihood(k, X, Y, 0.3), decimal=4)
    mu, var = m.predict_f(X + 1.0)
    mu_g, var_g = orc.gpr_predict_f(k, X, Y, 0.3, X + 1.0)
    np.testing.assert_array_almost_equal(mu, mu_g, decimal=4)
    np.testing.assert_array_almost_equal(var, var_g, decimal=4)
    e0 = m.elbo((X, Y))
    m.natgrad_step((X, Y), lr=0.9)   # :134-145 unchanged at the optimum
    np.testing.assert_almost_equal(e0, m.elbo((X, Y)), decimal=4)
    m.close()


def test_errors_leave_sites_unchanged():
    import tsvgp_b200 as tb
    rng = np.random.default_rng(0)
    X = rng.standard_normal((200, 2))
    Y = rng.standard_normal((200, 1))
    Z = np.concatenate([X[:10], X[:1]])   # duplicated inducing point: Kuu + 0*I is singular
    m = tb.t_SVGP(orc.SquaredExponential(lengthscales=1.0), orc.Gaussian(variance=0.1), Z)
    m.natgrad_step((X, Y), lr=0.5)         # fine with the default jitter
    l1, l2 = m.lambda_1, m.lambda_2_sqrt
    with pytest.raises(tb.InvalidArgumentError) as ei:
        m.natgrad_step((X, Y), lr=0.5, jitter=0.0)
    assert isinstance(ei.value, tb.NotPositiveDefiniteError) and ei.value.pivot == 11
    np.testing.assert_array_equal(m.lambda_1, l1)
    np.testing.assert_array_equal(m.lambda_2_sqrt, l2)
    with pytest.raises(tb.InvalidArgumentError):
        m.predict_f(np.zeros((3, 5)))     # wrong D
    with pytest.raises(tb.InvalidArgumentError):
        m.natgrad_step((np.zeros((0, 2)), np.zeros((0, 1))))     # empty minibatch
    with pytest.raises(tb.InvalidArgumentError):
        m.natgrad_step((X, Y[:-1]))       # ragged X / Y
    assert m.predict_f(np.zeros((0, 2)))[0].shape == (0, 1)
    np.testing.assert_array_equal(m.lambda_1, l1)
    m.natgrad_step((X, Y), lr=0.5)         # the context is still usable
    m.close()


def test_device_resident_dlpack_inputs():
    # tensors already on the GPU are aliased through DLPack (torch is only the producer here, not part of the product)
    torch = pytest.importorskip("torch")
    import tsvgp_b200 as tb
    import tsvgp_b200.synth as synth
    cfg = synth.describe("cfg3")
    X, Y, Z = synth.make_minibatch(cfg, n_rows=2000, M=128)
    kernel, lik = _objects(cfg)
    a = tb.t_SVGP(kernel, lik, Z.copy())
    b = tb.t_SVGP(kernel, lik, Z.copy())
    Xd, Yd = torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda()
    a.natgrad_step((X, Y), lr=0.5)
    b.set_data((Xd, Yd))
    b.natgrad_step(lr=0.5)
    np.testing.assert_array_equal(a.lambda_1, b.lambda_1)
    np.testing.assert_array_equal(a.lambda_2_sqrt, b.lambda_2_sqrt)
    a.close(); b.close()


def test_two_latent_gps_match_the_reference_fixture():
    # reference tests/models/test_tsvgp.py:45-88 runs its Bernoulli fixture with num_latent_gps in {1, 2}
    import tsvgp_b200 as tb
    rng = np.random.RandomState(123)
    X = rng.rand(40, 1) * 2 - 1
    F = np.stack([np.sin(3 * X[:, 0]), np.cos(4 * X[:, 0])], axis=1)
    Y = (F + 0.2 * rng.randn(40, 2) > 0).astype(float)
    kernel, lik = orc.SquaredExponential(lengthscales=0.15, variance=2.25), orc.Bernoulli()
    Z = np.linspace(-1, 1, 12)[:, None]      # cond(Kuu) ~ 1e2: the 1e-9 gate is meaningful
    ref = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(Z.copy()), num_latent_gps=2)
    dev = tb.t_SVGP(kernel, lik, Z.copy(), num_latent_gps=2)
    assert dev.lambda_1.shape == (12, 2) and dev.lambda_2_sqrt.shape == (2, 12, 12)
    for _ in range(3):
        ref.natgrad_step((X, Y), lr=0.8)
        dev.natgrad_step((X, Y), lr=0.8)
    errs = {"lambda_1": relerr(dev.lambda_1, ref.lambda_1), "lambda_2": relerr(dev.lambda_2, ref.lambda_2),
            "elbo": abs(dev.elbo((X, Y)) - ref.elbo((X, Y))) / abs(ref.elbo((X, Y)))}
    mu_d, var_d = dev.predict_f(X + 0.1)
    mu_r, var_r = ref.predict_f(X + 0.1)
    errs["mean"], errs["var"] = relerr(mu_d, mu_r), relerr(var_d, var_r)
    check(errs)
    dev.close()


@pytest.mark.parametrize("lengthscale", [1.0, 4.0, 8.0])
def test_cuda_is_as_close_to_exact_arithmetic_as_the_oracle(lengthscale):
    # the long-double arbiter (oracle/longdouble.py, reference operation order) is the truth; the float64 oracle sits
    # eps * cond away from it; the CUDA path must not sit further away than a small multiple of that
    import tsvgp_b200 as tb
    from oracle import longdouble as ld
    rng = np.random.RandomState(0)
    N, M, D = 150, 24, 8
    X, Z = rng.randn(N, D), rng.randn(M, D)
    Y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.randn(N, 1)
    kernel, lik = orc.SquaredExponential(variance=1.0, lengthscales=lengthscale), orc.Gaussian(variance=0.1)
    ref = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(Z))
    ref.natgrad_step((X, Y), lr=0.7)
    l1, L2 = ref.lambda_1.copy(), ref.lambda_2_sqrt.copy()
    t1, tL2 = ld.natgrad_step_gaussian(X, Y, Z, 1.0, lengthscale, 0.1, l1[:, 0], L2[0], lr=0.7)
    truth1, truth2 = np.asarray(t1, dtype=np.float64), np.asarray(tL2 @ tL2.T, dtype=np.float64)
    ref.natgrad_step((X, Y), lr=0.7)
    dev = tb.t_SVGP(kernel, lik, Z.copy(), lambda_1=l1, lambda_2_sqrt=L2)
    dev.natgrad_step((X, Y), lr=0.7)
    e_ref = max(relerr(ref.lambda_1[:, 0], truth1), relerr(ref.lambda_2[0], truth2))
    e_dev = max(relerr(dev.lambda_1[:, 0], truth1), relerr(dev.lambda_2[0], truth2))
    _record({"oracle_vs_longdouble": e_ref, "cuda_vs_longdouble": e_dev, "route": dev.timings()["route"], "cond_est": dev.timings()["cond_est"]})
    # fused route (cond below route_cond_max): eps * cond^2, required to stay 10x under the 1e-9 contract;
    # whitened route: the reference's own eps * cond
    assert e_dev <= max(20.0 * e_ref + 1e-13, 1e-10), (e_dev, e_ref)
    dev.close()


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg5"])
def test_predict_y_and_log_density(name):
    # the inherited GPModel surface used by the reference's experiment scripts (uci_regression.py:114,142,145,251)
    import tsvgp_b200 as tb
    import tsvgp_b200.synth as synth
    cfg = synth.describe(name)
    X, Y, Z = synth.make_minibatch(cfg, n_rows=1500, M=64)
    kernel, lik = synth.build_objects(cfg, orc)
    ref = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(Z.copy()))
    dev = tb.t_SVGP(kernel, lik, Z.copy())
    for _ in range(2):
        ref.natgrad_step((X, Y), lr=0.5)
        dev.natgrad_step((X, Y), lr=0.5)
    my_d, vy_d = dev.predict_y(X[:200])
    my_r, vy_r = orc.predict_y(ref, X[:200])
    errs = {"y_mean": relerr(my_d, my_r), "y_var": relerr(vy_d, vy_r),
            "log_density": relerr(dev.predict_log_density((X[:200], Y[:200])), orc.predict_log_density(ref, (X[:200], Y[:200])))}
    check(errs)
    assert abs(dev.training_loss_closure((X, Y))() + ref.elbo((X, Y))) <= 1e-9 * abs(ref.elbo((X, Y)))
    dev.close()


def test_prefetched_minibatch_stream_equals_direct_calls():
    # the input pipeline (stage_data / commit_staged / stream_minibatches) feeds exactly the same steps as passing data directly
    import tsvgp_b200 as tb
    import tsvgp_b200.synth as synth
    cfg = synth.describe("cfg3")
    kernel, lik = _objects(cfg)
    batches = []
    Z = None
    for i, n in enumerate([1500, 1500, 900, 2100]):          # ragged sizes: the staging buffers must re-grow
        X, Y, Zi = synth.make_minibatch(cfg, n_rows=n, M=160, seed_offset=i)
        Z = Zi if Z is None else Z
        px, py = tb.pinned_empty(X.shape), tb.pinned_empty(Y.shape)
        px[...] = X; py[...] = Y
        batches.append((px, py))
    a = tb.t_SVGP(kernel, lik, Z.copy(), num_data=60_000)
    b = tb.t_SVGP(kernel, lik, Z.copy(), num_data=60_000)
    for batch in batches:
        a.natgrad_step(batch, lr=0.5)
    rows = []
    for n in tb.stream_minibatches(b, batches):
        rows.append(n)
        b.natgrad_step(lr=0.5)
    assert rows == [1500, 1500, 900, 2100]
    np.testing.assert_array_equal(a.lambda_1, b.lambda_1)
    np.testing.assert_array_equal(a.lambda_2_sqrt, b.lambda_2_sqrt)
    for px, py in batches:
        tb.pinned_free(px); tb.pinned_free(py)
    a.close(); b.close()


@pytest.mark.parametrize("lik_name", ["bernoulli", "student_t"])
def test_cuda_quadrature_path_against_the_long_double_arbiter(lik_name):
    # the non-conjugate path (20-point Gauss-Hermite, analytic gradients, -1e-8 clip; tsvgp.py:256-263): the CUDA result must sit as
    # close to the 80-bit evaluation of the reference's formulas (oracle/longdouble.py::natgrad_step) as the float64 oracle does
    import tsvgp_b200 as tb
    from oracle import longdouble as ld
    rng = np.random.RandomState(1)
    N, M, D = 120, 20, 6
    X, Z = rng.randn(N, D), rng.randn(M, D)
    f = np.sin(X.sum(1, keepdims=True))
    kernel = orc.SquaredExponential(variance=1.0, lengthscales=2.0)
    if lik_name == "bernoulli":
        lik, spec, Y = orc.Bernoulli(), ("bernoulli",), (f + 0.3 * rng.randn(N, 1) > 0).astype(float)
    else:
        lik, spec, Y = orc.StudentT(scale=0.3, df=3.0), ("student_t", 0.3, 3.0), f + 0.3 * rng.standard_t(3.0, size=(N, 1))
    ref = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(Z), num_data=4 * N)
    ref.natgrad_step((X, Y), lr=0.6)
    l1, L2 = ref.lambda_1.copy(), ref.lambda_2_sqrt.copy()
    t1, tL2 = ld.natgrad_step(X, Y, Z, 1.0, 2.0, spec, l1[:, 0], L2[0], lr=0.6, scale=4.0)
    truth1, truth2 = np.asarray(t1, dtype=np.float64), np.asarray(tL2 @ tL2.T, dtype=np.float64)
    ref.natgrad_step((X, Y), lr=0.6)
    dev = tb.t_SVGP(kernel, lik, Z.copy(), lambda_1=l1, lambda_2_sqrt=L2, num_data=4 * N)
    dev.natgrad_step((X, Y), lr=0.6)
    e_ref = max(relerr(ref.lambda_1[:, 0], truth1), relerr(ref.lambda_2[0], truth2))
    e_dev = max(relerr(dev.lambda_1[:, 0], truth1), relerr(dev.lambda_2[0], truth2))
    _record({"oracle_vs_longdouble": e_ref, "cuda_vs_longdouble": e_dev, "route": dev.timings()["route"]})
    assert e_dev <= max(20.0 * e_ref + 1e-13, 1e-10), (e_dev, e_ref)
    dev.close()
