"""
GPflow's `SVGP` and `NaturalGradient` optimiser restated in NumPy.  TEST INFRASTRUCTURE ONLY (same rules as tsvgp_oracle.py:
only tests/ may import it; the product never does).

Why it exists: the reference's own tests pin the non-conjugate t-SVGP path and the M-step gradients ONLY by comparison with
GPflow's q-SVGP — `/root/reference/tests/models/test_tsvgp.py:45-88,123-131` (t-SVGP after 20 natural-gradient steps at lr = 1
predicts like `gpflow.models.SVGP` after 20 `NaturalGradient(gamma=1.0)` steps, Bernoulli likelihood, L in {1, 2}) and
`:168-188` (the gradients of -ELBO w.r.t. the kernel hyperparameters agree).  GPflow 2.2.1 / TensorFlow 2.5.0 are pinned by the
reference (`setup.py:3-8`) but are not under /root/reference and cannot be installed here, so their published algorithms are
restated [GPflow-recalled]:

  gpflow.models.SVGP (whiten=True is the constructor default, q_mu = 0, q_sqrt = I)
      predict_f  -> gpflow.conditionals.conditional(..., white=True) -> base_conditional
      prior_kl   -> gauss_kl(q_mu, q_sqrt, K=None)                    (whitened prior N(0, I))
      elbo       -> sum variational_expectations * scale - KL
  gpflow.optimizers.NaturalGradient(gamma) with the default XiNat transform (Salimbeni, Eleftheriadis & Hensman 2018):
      theta <- theta - gamma * dL/d eta,  theta = (S^-1 mu, -1/2 S^-1) natural, eta = (mu, S + mu mu^T) expectation parameters,
      through gpflow.optimizers.natgrad.{meanvarsqrt_to_natural, natural_to_meanvarsqrt} restated op for op.
GPflow obtains dL/d eta by TensorFlow autodiff through `expectation_to_meanvarsqrt`; here the same derivative is written in closed
form (dL/d eta1 = dL/d mu - 2 (dL/dS) mu, dL/d eta2 = dL/dS), with dL/d mu, dL/dS from the analytic Gauss-Hermite gradients the
t-SVGP oracle already uses (pinned by finite differences in tests/test_oracle_properties.py).
"""
from __future__ import annotations

import numpy as np

from . import tsvgp_oracle as orc


def _inverse_lower_triangular(L):
    """gpflow.optimizers.natgrad._inverse_lower_triangular: triangular_solve(L, I)."""
    return orc.triangular_solve(L, np.eye(L.shape[-1]))


def meanvarsqrt_to_natural(mu, s_sqrt):
    """gpflow.optimizers.natgrad.meanvarsqrt_to_natural, one latent: mu [M], s_sqrt [M, M] lower."""
    s_sqrt_inv = _inverse_lower_triangular(s_sqrt)
    s_inv = s_sqrt_inv.T @ s_sqrt_inv
    return s_inv @ mu, -0.5 * s_inv


def natural_to_meanvarsqrt(nat1, nat2):
    """gpflow.optimizers.natgrad.natural_to_meanvarsqrt (two Choleskys: S is needed as L L^T, not L^T L)."""
    var_sqrt_inv = orc.cholesky(-2.0 * nat2)
    var_sqrt = _inverse_lower_triangular(var_sqrt_inv)
    S = var_sqrt.T @ var_sqrt
    mu = S @ nat1
    return mu, orc.cholesky(S)


class OracleSVGP:
    """gpflow.models.SVGP(kernel, likelihood, inducing_variable, num_latent_gps=L) with its defaults: whiten=True, q_diag=False."""

    def __init__(self, kernel, likelihood, inducing_variable, *, num_latent_gps=1, whiten=True, num_data=None):
        self.kernel, self.likelihood = kernel, likelihood
        self.inducing_variable = inducing_variable if hasattr(inducing_variable, "Z") else orc.InducingPoints(inducing_variable)
        self.num_latent_gps, self.whiten, self.num_data = int(num_latent_gps), bool(whiten), num_data
        M = self.inducing_variable.num_inducing
        self.q_mu = np.zeros((M, self.num_latent_gps))                                  # SVGP._init_variational_parameters
        self.q_sqrt = np.stack([np.eye(M) for _ in range(self.num_latent_gps)])

    # -- conditional ------------------------------------------------------------------------------------------------------------
    def _projection(self, Xnew):
        """A [M, N] with f_mean = A^T q_mu: Lm^-1 Kmn when whitened, Lm^-T Lm^-1 Kmn otherwise; and the prior part of f_var."""
        Kmm = orc.Kuu(self.inducing_variable, self.kernel, jitter=orc.DEFAULT_JITTER)
        Kmn = orc.Kuf(self.inducing_variable, self.kernel, np.asarray(Xnew, dtype=np.float64))
        Lm = orc.cholesky(Kmm)
        A = orc.triangular_solve(Lm, Kmn)
        fvar0 = self.kernel.K_diag(Xnew) - np.sum(np.square(A), axis=0)
        if not self.whiten:
            A = orc.triangular_solve(Lm, A, adjoint=True)
        return A, fvar0, Lm

    def predict_f(self, Xnew):
        A, fvar0, _ = self._projection(Xnew)
        fmean = A.T @ self.q_mu
        fvar = np.stack([fvar0 + np.sum(np.square(np.tril(self.q_sqrt[l]).T @ A), axis=0) for l in range(self.num_latent_gps)], axis=1)
        return fmean, fvar

    def prior_kl(self):
        if self.whiten:   # gauss_kl(q_mu, q_sqrt, None)
            M, L = self.q_mu.shape
            Lq = np.tril(self.q_sqrt)
            logdet_qcov = np.sum(np.log(np.square(np.diagonal(Lq, axis1=-2, axis2=-1))))
            return 0.5 * (np.sum(np.square(self.q_mu)) - M * L - logdet_qcov + np.sum(np.square(Lq)))
        return orc.gauss_kl(self.q_mu, self.q_sqrt, orc.Kuu(self.inducing_variable, self.kernel, jitter=orc.DEFAULT_JITTER))

    def _scale(self, n):
        return self.num_data / n if self.num_data is not None else 1.0

    def elbo(self, data):
        X, Y = np.asarray(data[0], dtype=np.float64), np.asarray(data[1], dtype=np.float64)
        fmean, fvar = self.predict_f(X)
        ve = self.likelihood.variational_expectations(fmean, fvar, Y)
        return np.sum(ve) * self._scale(X.shape[0]) - self.prior_kl()

    def training_loss(self, data):
        return -self.elbo(data)

    # -- NaturalGradient(gamma).minimize(training_loss, var_list=[(q_mu, q_sqrt)]) ------------------------------------------
    def natgrad_step(self, data, gamma=1.0):
        X, Y = np.asarray(data[0], dtype=np.float64), np.asarray(data[1], dtype=np.float64)
        A, fvar0, Lm = self._projection(X)
        fmean, fvar = self.predict_f(X)
        _, g_m, g_v = self.likelihood.ve_and_grads(fmean, fvar, Y)     # d sum(ve) / d fmean, / d fvar   [N, L]
        s = self._scale(X.shape[0])
        M = self.q_mu.shape[0]
        new_mu, new_sqrt = np.empty_like(self.q_mu), np.empty_like(self.q_sqrt)
        Kinv = None
        if not self.whiten:
            Kinv = orc.cholesky_solve(Lm, np.eye(M))
        for l in range(self.num_latent_gps):
            mu, Lq = self.q_mu[:, l], np.tril(self.q_sqrt[l])
            nat1, nat2 = meanvarsqrt_to_natural(mu, Lq)
            S_inv = -2.0 * nat2
            P = np.eye(M) if self.whiten else Kinv                      # prior precision
            dL_dmu = -s * (A @ g_m[:, l]) + P @ mu                      # L = -ELBO
            dL_dS = -s * ((A * g_v[:, l]) @ A.T) + 0.5 * (P - S_inv)
            dL_deta1 = dL_dmu - 2.0 * dL_dS @ mu
            dL_deta2 = dL_dS
            new_mu[:, l], new_sqrt[l] = natural_to_meanvarsqrt(nat1 - gamma * dL_deta1, nat2 - gamma * dL_deta2)
        self.q_mu, self.q_sqrt = new_mu, new_sqrt
