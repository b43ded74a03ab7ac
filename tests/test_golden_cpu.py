"""The oracle reproduces the committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import tsvgp_oracle as orc
import tsvgp_b200.synth as synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]


def relerr(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(name):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    cfg = synth.describe(name)
    # the fixture's inputs are what the synthetic generator makes (guards the generator too)
    X, Y, Z = synth.make_minibatch(cfg, n_rows=g["X"].shape[0], M=g["Z"].shape[0])
    np.testing.assert_array_equal(X, g["X"]); np.testing.assert_array_equal(Y, g["Y"]); np.testing.assert_array_equal(Z, g["Z"])
    kernel, lik = synth.build_objects(cfg, orc)
    nd = None if int(g["num_data"]) < 0 else int(g["num_data"])
    m = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(Z.copy()), num_data=nd)
    tol = 1e-10   # same algorithm, possibly another BLAS / thread count
    for s in range(2):
        assert abs(m.elbo((X, Y)) - g[f"elbo_before_{s}"]) <= tol * abs(g[f"elbo_before_{s}"])
        m.natgrad_step((X, Y), lr=float(g["lr"]))
        assert relerr(m.lambda_1, g[f"lambda_1_{s}"]) <= tol
        assert relerr(m.lambda_2, g[f"lambda_2_{s}"]) <= tol
    mu, var = m.predict_f(g["Xt"])
    assert relerr(mu, g["mean"]) <= tol and relerr(var, g["var"]) <= tol
    assert abs(m.prior_kl() - g["prior_kl"]) <= tol * abs(g["prior_kl"])
