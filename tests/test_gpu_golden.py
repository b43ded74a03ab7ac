"""GPU: the CUDA path against the committed golden vectors (tests/golden/*.npz) — no oracle code runs here."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-9


def relerr(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
def test_cuda_reproduces_golden(name):
    import tsvgp_b200 as tb
    import tsvgp_b200.synth as synth
    from tsvgp_b200 import standins as st
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    cfg = synth.describe(name)
    kernel, lik = synth.build_objects(cfg, st)
    nd = None if int(g["num_data"]) < 0 else int(g["num_data"])
    m = tb.t_SVGP(kernel, lik, st.InducingPoints(g["Z"]), num_data=nd)
    X, Y = g["X"], g["Y"]
    for s in range(2):
        e = m.natgrad_step((X, Y), lr=float(g["lr"]), return_elbo=True)
        assert abs(e - g[f"elbo_before_{s}"]) <= TOL * abs(g[f"elbo_before_{s}"])
        assert relerr(m.lambda_1, g[f"lambda_1_{s}"]) <= TOL
        assert relerr(m.lambda_2, g[f"lambda_2_{s}"]) <= TOL
    mu, var = m.predict_f(g["Xt"])
    assert relerr(mu, g["mean"]) <= TOL and relerr(var, g["var"]) <= TOL
    assert abs(m.elbo((X, Y)) - g["elbo_after"]) <= TOL * abs(g["elbo_after"])
    assert abs(m.prior_kl() - g["prior_kl"]) <= TOL * abs(g["prior_kl"])
    m.close()
