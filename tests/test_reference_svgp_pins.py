"""
The reference's remaining pins of the hot path, restated on the CPU (VERDICT r01 "what's missing" #2):

  /root/reference/tests/models/test_tsvgp.py:45-88,123-131   t-SVGP after 20 natgrad steps at lr = 1 == GPflow SVGP after 20
                                                             NaturalGradient(gamma=1.0) steps; Bernoulli; L in {1, 2}; decimal=4
  /root/reference/tests/models/test_tsvgp.py:168-188         d(-ELBO)/d(kernel hyperparameters) of the two models agree; decimal=4

These are the only reference-held checks of the non-conjugate (Gauss-Hermite) natural-gradient path and of the M-step
gradients.  GPflow cannot run here, so SVGP + NaturalGradient are restated in oracle/svgp_oracle.py; what this file pins is the
t-SVGP oracle (oracle/tsvgp_oracle.py: natgrad_step, new_predict_f, predict_f, elbo, elbo_gradients) against that independent
route to the same posterior — two different parameterisations, update rules and conditionals that only agree if both are right.
"""
import numpy as np
import pytest

from oracle import svgp_oracle as sv
from oracle import tsvgp_oracle as orc

LENGTH_SCALE, VARIANCE, NUM_DATA, NOISE_VARIANCE = 2.0, 2.25, 8, 0.3     # test_tsvgp.py:12-15


def _setup(rng):
    """test_tsvgp.py:91-103"""
    def func(x):
        return np.sin(x * 3 * 3.14) + 0.3 * np.cos(x * 9 * 3.14) + 0.5 * np.sin(x * 7 * 3.14)

    X = rng.rand(NUM_DATA, 1) * 2 - 1
    Y = func(X) + 0.2 * rng.randn(NUM_DATA, 1)
    return X, Y, orc.SquaredExponential(lengthscales=LENGTH_SCALE, variance=VARIANCE)


def _tsvgp_qsvgp_optim_setup(num_latent_gps, binary_labels=False, steps=20):
    """test_tsvgp.py:45-88.  The reference multiplies the 0/1 labels by np.random.rand(1, L) (unseeded global RNG), so every label
    differs from 1 and GPflow's Bernoulli treats it as class 0; `binary_labels=True` keeps proper 0/1 labels as a second case."""
    rng = np.random.RandomState(123)
    X, obs, kernel = _setup(rng)
    Y = np.tile((obs > 0.0).astype(float), [1, num_latent_gps])
    if not binary_labels:
        Y = Y * np.random.RandomState(7).rand(1, num_latent_gps)
    lik = orc.Bernoulli()
    svgp = sv.OracleSVGP(kernel, lik, orc.InducingPoints(X.copy()), num_latent_gps=num_latent_gps)
    tsvgp = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(X.copy()), num_latent_gps=num_latent_gps)
    for _ in range(steps):
        tsvgp.natgrad_step((X, Y), lr=1.0)
    for _ in range(steps):
        svgp.natgrad_step((X, Y), gamma=1.0)
    return tsvgp, svgp, (X, Y)


@pytest.mark.parametrize("binary_labels", [False, True])
@pytest.mark.parametrize("num_latent_gps", [1, 2])
def test_predictions_match_tsvgp_qsvgp_optimal(num_latent_gps, binary_labels):
    # test_tsvgp.py:123-131
    tsvgp, qsvgp, data = _tsvgp_qsvgp_optim_setup(num_latent_gps, binary_labels)
    X = data[0] + 0.1
    mu_t, var_t = tsvgp.new_predict_f(X)
    mu_q, var_q = qsvgp.predict_f(X)
    np.testing.assert_array_almost_equal(mu_t, mu_q, decimal=4)
    np.testing.assert_array_almost_equal(var_t, var_q, decimal=4)
    # the path's own conditional (predict_f, tsvgp.py:97-114) gives the same moments as the alternative algebra
    mu_p, var_p = tsvgp.predict_f(X)
    np.testing.assert_array_almost_equal(mu_p, mu_q, decimal=4)
    np.testing.assert_array_almost_equal(var_p, var_q, decimal=4)
    # matched posteriors: same bound
    np.testing.assert_almost_equal(tsvgp.elbo(data), qsvgp.elbo(data), decimal=4)


def test_one_step_each_is_already_the_same_posterior_for_a_gaussian_likelihood():
    # conjugate case: one natural-gradient step at lr = gamma = 1 from the prior lands both models on the optimum
    rng = np.random.RandomState(123)
    X, Y, kernel = _setup(rng)
    lik = orc.Gaussian(variance=NOISE_VARIANCE)
    svgp = sv.OracleSVGP(kernel, lik, orc.InducingPoints(X.copy()))
    tsvgp = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(X.copy()))
    tsvgp.natgrad_step((X, Y), lr=1.0)
    svgp.natgrad_step((X, Y), gamma=1.0)
    mu_t, var_t = tsvgp.predict_f(X + 0.1)
    mu_q, var_q = svgp.predict_f(X + 0.1)
    np.testing.assert_array_almost_equal(mu_t, mu_q, decimal=4)
    np.testing.assert_array_almost_equal(var_t, var_q, decimal=4)
    np.testing.assert_almost_equal(svgp.elbo((X, Y)), orc.gpr_log_marginal_likelihood(kernel, X, Y, NOISE_VARIANCE), decimal=4)


def _softplus_jacobian(theta):
    """d theta / d u for GPflow's default positive transform theta = softplus(u): the reference differentiates w.r.t. the
    unconstrained `trainable_variables`."""
    return 1.0 - np.exp(-theta)


def _svgp_loss_grads(svgp, data, rel_step=1e-5):
    """d training_loss / d (unconstrained variance, lengthscales) of the SVGP by central differences, q(u) held fixed."""
    out = []
    for name in ("variance", "lengthscales"):
        base = float(getattr(svgp.kernel, name))
        h = rel_step * base
        vals = []
        for sgn in (+1.0, -1.0):
            setattr(svgp.kernel, name, orc._param(base + sgn * h))
            vals.append(svgp.training_loss(data))
        setattr(svgp.kernel, name, orc._param(base))
        out.append((vals[0] - vals[1]) / (2 * h) * _softplus_jacobian(base))
    return np.array(out)


@pytest.mark.parametrize("num_latent_gps", [1, 2])
def test_gradient_wrt_hyperparameters(num_latent_gps):
    # test_tsvgp.py:168-188: for matched posteriors the gradients of -ELBO w.r.t. the kernel hyperparameters agree.
    # GPflow orders kernel.trainable_variables as (variance, lengthscales) for a Stationary kernel.
    tsvgp, qsvgp, data = _tsvgp_qsvgp_optim_setup(num_latent_gps)
    grads_q = _svgp_loss_grads(qsvgp, data)
    # analytic d ELBO / d (constrained) hyperparameters with the sites fixed; the ELBO is additive over the latent GPs (shared
    # kernel: variational_expectations sums over the latent axis and the KL over the L independent q(u_l), tsvgp.py:65-95)
    d_var = d_ls = 0.0
    for l in range(num_latent_gps):
        one = orc.OracleTSVGP(tsvgp.kernel, tsvgp.likelihood, tsvgp.inducing_variable, lambda_1=tsvgp.lambda_1[:, l:l + 1],
                              lambda_2_sqrt=tsvgp.lambda_2_sqrt[l:l + 1])
        _, g = orc.elbo_gradients(one, (data[0], data[1][:, l:l + 1]))
        d_var, d_ls = d_var + g["variance"], d_ls + float(np.sum(g["lengthscales"]))
    grads_t = -np.array([d_var * _softplus_jacobian(VARIANCE), d_ls * _softplus_jacobian(LENGTH_SCALE)])
    np.testing.assert_array_almost_equal(grads_q, grads_t, decimal=4)
    # and the analytic t-SVGP gradient is the derivative of the reference-order ELBO (central differences, sites held fixed)
    fd = []
    for name in ("variance", "lengthscales"):
        base = float(getattr(tsvgp.kernel, name))
        h = 1e-5 * base
        vals = []
        for sgn in (+1.0, -1.0):
            setattr(tsvgp.kernel, name, orc._param(base + sgn * h))
            vals.append(-tsvgp.elbo(data))
        setattr(tsvgp.kernel, name, orc._param(base))
        fd.append((vals[0] - vals[1]) / (2 * h) * _softplus_jacobian(base))
    np.testing.assert_array_almost_equal(np.array(fd), grads_t, decimal=4)
