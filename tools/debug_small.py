"""Tooling: one small natgrad_step/elbo/predict_f (for compute-sanitizer runs).  usage: debug_small.py N M D [lik]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tsvgp_b200 as tb
from tsvgp_b200 import standins as st

N, M, D = (int(a) for a in sys.argv[1:4])
lik = sys.argv[4] if len(sys.argv) > 4 else "Gaussian"
rng = np.random.default_rng(0)
X = rng.standard_normal((N, D)); Y = np.sin(X.sum(1, keepdims=True)) + 0.3 * rng.standard_normal((N, 1))
if lik == "Bernoulli":
    Y = (Y > 0).astype(float)
likelihood = {"Gaussian": st.Gaussian(0.1), "Bernoulli": st.Bernoulli(), "StudentT": st.StudentT(0.3, 3.0)}[lik]
m = tb.t_SVGP(st.Matern52(1.0, 2.0), likelihood, X[:M].copy())
for i in range(2):
    print("elbo before", m.natgrad_step((X, Y), lr=0.5, return_elbo=True), m.timings())
print("elbo", m.elbo((X, Y)))
mu, var = m.predict_f(X[:10])
print(mu[:3, 0], var[:3, 0])
print("kl", m.prior_kl())
m.get_mean_chol_cov_inducing_posterior()
print("ok")
