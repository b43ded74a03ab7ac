"""
Generates tests/golden/*.npz: frozen inputs and outputs of the reference-order oracle (oracle/tsvgp_oracle.py) for reduced
instances of the five BASELINE.json configs.  The reference itself (GPflow 2.2.1 / TensorFlow 2.5.0) cannot be imported in
this image, so these vectors pin the ORACLE (and through it the CUDA path) against regressions; they are not outputs of
TensorFlow.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import tsvgp_oracle as orc  # noqa: E402
import tsvgp_b200.synth as synth  # noqa: E402

CASES = {  # name -> (config, rows, M, num_data, steps)
    "cfg1": ("cfg1", 2000, 50, None, 2),
    "cfg2": ("cfg2", 1500, 96, 15_000, 2),
    "cfg3": ("cfg3", 1200, 160, 12_000, 2),
    "cfg4": ("cfg4", 1000, 130, None, 2),
    "cfg5": ("cfg5", 1000, 128, 25_000, 2),
}


def run_case(name):
    cfg_name, n, M, num_data, steps = CASES[name]
    cfg = synth.describe(cfg_name)
    X, Y, Z = synth.make_minibatch(cfg, n_rows=n, M=M)
    kernel, lik = synth.build_objects(cfg, orc)
    m = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(Z.copy()), num_data=num_data)
    out = dict(X=X, Y=Y, Z=Z, lr=cfg["lr"], num_data=-1 if num_data is None else num_data)
    for s in range(steps):
        out[f"elbo_before_{s}"] = m.elbo((X, Y))
        m.natgrad_step((X, Y), lr=cfg["lr"])
        out[f"lambda_1_{s}"] = m.lambda_1.copy()
        out[f"lambda_2_{s}"] = m.lambda_2.copy()
    Xt = X[:64] + 0.05
    mu, var = m.predict_f(Xt)
    out.update(Xt=Xt, mean=mu, var=var, elbo_after=m.elbo((X, Y)), prior_kl=m.prior_kl())
    return out


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name in CASES:
        np.savez_compressed(os.path.join(here, f"{name}.npz"), **run_case(name))
        print("wrote", name)
