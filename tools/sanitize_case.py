"""Tooling: a small tour of the device paths for `compute-sanitizer --tool memcheck python tools/sanitize_case.py`
(one tool per run, smallest cases that still reach every kernel family: look-ahead Cholesky with strip kernels at 5 tile rows,
multi-slab pass on two streams, whitened and exact routes, several latents, Softmax Monte Carlo, M-step gradients, whitened sibling)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tsvgp_b200 as tb  # noqa: E402
from tsvgp_b200 import standins as st  # noqa: E402

rng = np.random.default_rng(0)


def data(n, d, m):
    X = rng.standard_normal((n, d))
    return X, X[:m].copy()


# Gaussian, M = 600 (5 tile rows: look-ahead Cholesky + strip kernels), 8 slabs, early slabs on the second step
X, Z = data(2100, 4, 600)
Y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((2100, 1))
m = tb.t_SVGP(st.Matern52(variance=1.0, lengthscales=2.0), st.Gaussian(variance=0.1), Z, num_data=21000)
m.set_option("chunk", 256)
for _ in range(3):
    m.set_option("invalidate", 1)
    e = m.natgrad_step((X, Y), lr=0.5, return_elbo=True)
print("gaussian", e, m.elbo(), m.predict_f(X[:33])[1].min(), m.elbo_and_grad()[1]["variance"])
for route in (2, 3):
    m.set_option("route", route)
    m.natgrad_step((X, Y), lr=0.5)
m.close()

# Bernoulli, two latents, ragged sizes
X, Z = data(777, 3, 130)
Y = (np.stack([np.sin(X[:, 0]), np.cos(X[:, 1])], 1) + 0.2 * rng.standard_normal((777, 2)) > 0).astype(float)
m = tb.t_SVGP(st.SquaredExponential(variance=1.0, lengthscales=1.1), st.Bernoulli(), Z, num_latent_gps=2)
m.set_option("chunk", 256)
for _ in range(2):
    m.natgrad_step((X, Y), lr=0.6)
print("bernoulli L=2", m.elbo((X, Y)), m.predict_f(X[:5])[0].shape, m.get_mean_chol_cov_inducing_posterior()[1].shape, m.elbo_and_grad((X, Y))[0])
m.close()


# Softmax, 3 classes, generator and explicit draws
X, Z = data(500, 2, 40)
lab = rng.integers(0, 3, (500, 1)).astype(float)
m = tb.t_SVGP(st.SquaredExponential(variance=1.0, lengthscales=0.9), st.Softmax(3, 7), Z, num_latent_gps=3)
m.natgrad_step((X, lab), lr=0.3)
m.set_mc_epsilon(rng.standard_normal((7, 500, 3)))
m.natgrad_step((X, lab), lr=0.3)
print("softmax", m.elbo((X, lab)), m.predict_f(X[:4])[0].shape)
m.close()

# whitened sibling + Student-t
X, Z = data(900, 3, 150)
Y = np.sin(X.sum(1, keepdims=True)) + 0.3 * rng.standard_t(3.0, size=(900, 1))
w = tb.t_SVGP_white(st.SquaredExponential(variance=1.0, lengthscales=1.3), st.Gaussian(variance=0.2), Z)
w.natgrad_step((X, Y), lr=0.7)
print("white", w.elbo((X, Y)), w.predict_f_extra_data(X[:9], (X[100:300], Y[100:300]))[0].shape)
w.close()
w = tb.t_SVGP_white(st.SquaredExponential(variance=1.0, lengthscales=1.3), st.Gaussian(variance=0.2), Z, num_latent_gps=2)
w.natgrad_step((X, np.hstack([Y, -Y])), lr=0.7)
print("white L=2", w.elbo(), w.predict_f_extra_data(X[:9], (X[100:300], np.hstack([Y, -Y])[100:300]))[0].shape, w.lambda_2.shape)
w.close()
s = tb.t_SVGP(st.SquaredExponential(variance=1.0, lengthscales=1.3), st.StudentT(scale=0.3, df=3.0), Z, num_data=9000)
s.natgrad_step((X, Y), lr=0.3)
print("student-t", s.elbo((X, Y)))
s.close()
print("SANITIZE_CASE_DONE")
