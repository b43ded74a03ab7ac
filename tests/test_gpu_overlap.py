"""
GPU: the step's overlap machinery gives the numbers of the plain schedule.
  * speculative route: the Kuu + jitter I chain runs underneath the streaming pass and the conditioning probe is read after it
    (tsvgp.cu::tsvgp_natgrad_step); a wrong guess repeats the pass with the right route.
  * early slabs: Gaussian likelihood, fused route — Kuf and the constant-weight SYRK of the first slab of every slab stream are
    enqueued before the posterior chain; their means come from a mat-vec over the finished slab and b += Kuf g from its own kernel.
Reference being matched: src/models/tsvgp.py:234-304 through the oracle, tolerance 1e-9 (norm-wise).
"""
import numpy as np
import pytest

from oracle import tsvgp_oracle as orc
from tests.test_gpu_parity import check, relerr

pytestmark = pytest.mark.gpu


def test_early_slabs_and_speculation_equal_the_plain_schedule_and_the_oracle():
    import tsvgp_b200 as tb
    import tsvgp_b200.synth as synth
    cfg = synth.describe("cfg3")                         # Gaussian, Matern-5/2
    X, Y, Z = synth.make_minibatch(cfg, n_rows=4096, M=1536)   # 12 tile rows: 78 SYRK tiles, no split-K -> early slabs eligible
    kernel, lik = synth.build_objects(cfg, orc)
    fast = tb.t_SVGP(kernel, lik, Z.copy(), num_data=40_960)
    plain = tb.t_SVGP(kernel, lik, Z.copy(), num_data=40_960)
    ref = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(Z.copy()), num_data=40_960)
    for m in (fast, plain):
        m.set_option("chunk", 512)                       # 8 slabs over 2 streams
    plain.set_option("speculate", 0); plain.set_option("early_slabs", 0)
    errs = {}
    for s in range(3):
        for m in (fast, plain):
            m.set_option("invalidate", 1)                # the K9 chain runs every step: steps 1, 2 speculate on step 0's estimate
        e_f = fast.natgrad_step((X, Y), lr=0.5, return_elbo=True)
        e_p = plain.natgrad_step((X, Y), lr=0.5, return_elbo=True)
        e_r = ref.elbo((X, Y))
        ref.natgrad_step((X, Y), lr=0.5)
        assert fast.timings()["route"] == 1 and plain.timings()["route"] == 1
        assert abs(e_f - e_p) <= 1e-12 * abs(e_p)
        assert relerr(fast.lambda_1, plain.lambda_1) < 1e-11 and relerr(fast.lambda_2, plain.lambda_2) < 1e-11
        errs[f"elbo_before_step{s}"] = abs(e_f - e_r) / abs(e_r)
        errs[f"lambda_1_step{s}"] = relerr(fast.lambda_1, ref.lambda_1)
        errs[f"lambda_2_step{s}"] = relerr(fast.lambda_2, ref.lambda_2)
    check(errs)
    fast.close(); plain.close()


def test_a_wrong_route_guess_repeats_the_pass():
    # step 0 at a short lengthscale (cond ~ 3e3: fused); then the kernel changes to a long one (cond ~ 1e7: whitened).  The next step
    # speculates "fused" from the stale estimate, reads the probe after the pass and must repeat it on the whitened route.
    import tsvgp_b200 as tb
    import tsvgp_b200.synth as synth
    cfg = synth.describe("cfg2", ls=1.414)
    X, Y, Z = synth.make_minibatch(cfg, n_rows=2000, M=300)
    kernel, lik = synth.build_objects(cfg, orc)
    dev = tb.t_SVGP(kernel, lik, Z.copy())
    ref = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(Z.copy()))
    dev.natgrad_step((X, Y), lr=0.5); ref.natgrad_step((X, Y), lr=0.5)
    assert dev.timings()["route"] == 1
    kernel.lengthscales = orc._param(4.0)
    dev.natgrad_step((X, Y), lr=0.5); ref.natgrad_step((X, Y), lr=0.5)
    assert dev.timings()["route"] == 2 and dev.timings()["cond_est"] > 1e4
    # parity at the reference's own rounding level for this conditioning (tests/test_gpu_parity.py, ill-conditioned case)
    check({"lambda_1": relerr(dev.lambda_1, ref.lambda_1), "lambda_2": relerr(dev.lambda_2, ref.lambda_2)}, tol=1e-6)
    dev.close()


def test_route_option_is_validated():
    import tsvgp_b200 as tb
    m = tb.t_SVGP(orc.SquaredExponential(), orc.Gaussian(), np.zeros((3, 1)) + np.arange(3)[:, None])
    with pytest.raises(tb.InvalidArgumentError):
        m.set_option("route", 4)
    m.close()
