"""
CPU oracle for the t-SVGP hot path.  TEST INFRASTRUCTURE ONLY.

This is a float64 NumPy/SciPy restatement, op for op and in the reference's own
operation order, of

  * ``t_SVGP.natgrad_step``                      /root/reference/src/models/tsvgp.py:234-304
  * ``base_SVGP.elbo`` / ``prior_kl``            /root/reference/src/models/tsvgp.py:65-95
  * ``base_SVGP.predict_f``                      /root/reference/src/models/tsvgp.py:97-114
  * ``t_SVGP.get_mean_chol_cov_inducing_posterior``   tsvgp.py:202-212
  * ``t_SVGP.new_predict_f``                     tsvgp.py:215-232
  * ``posterior_from_dense_site``                /root/reference/src/util.py:349-391
  * ``conditional_from_precision_sites``         /root/reference/src/util.py:91-185
  * ``gradient_transformation_mean_var_to_expectation``   util.py:429-438
  * ``DenseSites``                               /root/reference/src/sites.py:43-80

The arithmetic the reference delegates to its pinned third-party dependencies
(gpflow==2.2.1, tensorflow==2.5.0, tensorflow-probability==0.13.0; reference
setup.py:3-8) is NOT under /root/reference and cannot be installed here (no wheels,
Python 3.12, no network).  Their published algorithms are restated below (SURVEY.md
Appendix B) and each function names the GPflow module it restates.

PARITY STATUS: **unpinned at 1e-9**.  The reference holds no golden vectors; its own
tests pin this path only by properties to 4 decimals (t-SVGP == exact GP regression at
the optimum, etc.; reference tests/models/test_tsvgp.py:106-165).  Those properties are
reproduced against closed-form GP regression in tests/test_oracle_properties.py; bit-level
agreement with GPflow/TensorFlow cannot be verified in this container.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module.  The product path (``t-svgp_b200``) never does.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg as sla
from scipy.special import erf, gammaln

DEFAULT_JITTER = 1e-6  # gpflow.config.default_jitter() in GPflow 2.2.1
N_GH = 20  # gpflow.likelihoods.ScalarLikelihood default number of Gauss-Hermite points


# ----------------------------------------------------------------------------------------
# Minimal stand-ins for the GPflow objects the model reads (attribute names as in GPflow).
# ----------------------------------------------------------------------------------------
class _Value(float):
    """float that also answers ``.numpy()`` like a gpflow.Parameter."""

    def numpy(self):
        return float(self)


class _ArrayValue(np.ndarray):
    def numpy(self):
        return np.asarray(self)


def _param(v):
    a = np.asarray(v, dtype=np.float64)
    if a.ndim == 0:
        return _Value(float(a))
    return a.view(_ArrayValue)


class SquaredExponential:
    """gpflow.kernels.SquaredExponential (stationaries.py): K = variance * exp(-r2 / 2)."""

    name = "squared_exponential"

    def __init__(self, variance=1.0, lengthscales=1.0):
        self.variance = _param(variance)
        self.lengthscales = _param(lengthscales)

    def K_r2(self, r2):
        return float(self.variance) * np.exp(-0.5 * r2)

    def K(self, X, X2=None):
        return self.K_r2(square_distance(scale(X, self.lengthscales), None if X2 is None else scale(X2, self.lengthscales)))

    def K_diag(self, X):
        return np.full(X.shape[0], float(self.variance))


RBF = SquaredExponential


class Matern52:
    """gpflow.kernels.Matern52: r = sqrt(max(r2, 1e-36)); variance*(1+sqrt5 r+5/3 r^2) exp(-sqrt5 r)."""

    name = "matern52"

    def __init__(self, variance=1.0, lengthscales=1.0):
        self.variance = _param(variance)
        self.lengthscales = _param(lengthscales)

    def K_r2(self, r2):
        r = np.sqrt(np.maximum(r2, 1e-36))
        sqrt5 = np.sqrt(5.0)
        return float(self.variance) * (1.0 + sqrt5 * r + 5.0 / 3.0 * np.square(r)) * np.exp(-sqrt5 * r)

    def K(self, X, X2=None):
        return self.K_r2(square_distance(scale(X, self.lengthscales), None if X2 is None else scale(X2, self.lengthscales)))

    def K_diag(self, X):
        return np.full(X.shape[0], float(self.variance))


def scale(X, lengthscales):
    """gpflow.kernels.Stationary.scale: X / lengthscales (broadcast; ARD allowed)."""
    return np.asarray(X, dtype=np.float64) / np.asarray(lengthscales, dtype=np.float64)


def square_distance(X, X2):
    """gpflow.utilities.ops.square_distance — expansion form, may go slightly negative."""
    if X2 is None:
        Xs = np.sum(np.square(X), axis=-1, keepdims=True)
        dist = -2.0 * (X @ X.T)
        dist += Xs + Xs.T
        return dist
    Xs = np.sum(np.square(X), axis=-1)
    X2s = np.sum(np.square(X2), axis=-1)
    dist = -2.0 * (X @ X2.T)
    dist += Xs[:, None] + X2s[None, :]
    return dist


class InducingPoints:
    """gpflow.inducing_variables.InducingPoints."""

    def __init__(self, Z):
        self.Z = np.array(Z, dtype=np.float64).view(_ArrayValue)

    @property
    def num_inducing(self):
        return self.Z.shape[0]


def Kuu(iv, kernel, jitter=0.0):
    """gpflow.covariances.Kuu(InducingPoints, Kernel): kernel(Z) + jitter*I."""
    Z = np.asarray(iv.Z)
    return kernel.K(Z) + jitter * np.eye(Z.shape[0])


def Kuf(iv, kernel, Xnew):
    """gpflow.covariances.Kuf(InducingPoints, Kernel, Xnew): kernel(Z, Xnew) -> [M, N]."""
    return kernel.K(np.asarray(iv.Z), np.asarray(Xnew, dtype=np.float64))


# ----------------------------------------------------------------------------------------
# Likelihoods (gpflow.likelihoods): variational_expectations and its (mean, var) gradients.
# The reference obtains the gradients by tf.GradientTape (tsvgp.py:256-259); the analytic
# derivative of the same closed form / quadrature sum is the same function.
# ----------------------------------------------------------------------------------------
def gh_points_and_weights(n_gh=N_GH):
    """gpflow.quadrature.gauss_hermite: nodes sqrt(2)*x_k, weights w_k/sqrt(pi)."""
    x, w = np.polynomial.hermite.hermgauss(n_gh)
    return x * np.sqrt(2.0), w / np.sqrt(np.pi)


class Gaussian:
    """gpflow.likelihoods.Gaussian."""

    name = "gaussian"

    def __init__(self, variance=1.0):
        self.variance = _param(variance)

    def variational_expectations(self, Fmu, Fvar, Y):
        s2 = float(self.variance)
        return np.sum(-0.5 * np.log(2 * np.pi) - 0.5 * np.log(s2) - 0.5 * (np.square(Y - Fmu) + Fvar) / s2, axis=-1)

    def ve_and_grads(self, Fmu, Fvar, Y):
        s2 = float(self.variance)
        ve = self.variational_expectations(Fmu, Fvar, Y)
        return ve, (Y - Fmu) / s2, np.full_like(Fvar, -0.5 / s2)

    def predict_mean_and_var(self, Fmu, Fvar):
        return Fmu, Fvar + float(self.variance)


class _QuadratureLikelihood:
    n_gh = N_GH

    def _logp(self, F, Y):
        raise NotImplementedError

    def _dlogp(self, F, Y):
        raise NotImplementedError

    def variational_expectations(self, Fmu, Fvar, Y):
        z, w = gh_points_and_weights(self.n_gh)
        F = Fmu[..., None] + np.sqrt(Fvar)[..., None] * z
        return np.sum(np.sum(self._logp(F, Y[..., None]) * w, axis=-1), axis=-1)

    def ve_and_grads(self, Fmu, Fvar, Y):
        z, w = gh_points_and_weights(self.n_gh)
        sd = np.sqrt(Fvar)
        F = Fmu[..., None] + sd[..., None] * z
        ve = np.sum(np.sum(self._logp(F, Y[..., None]) * w, axis=-1), axis=-1)
        d = self._dlogp(F, Y[..., None]) * w
        g_mean = np.sum(d, axis=-1)
        g_var = np.sum(d * z, axis=-1) / (2.0 * sd)
        return ve, g_mean, g_var


def inv_probit(x):
    """gpflow.likelihoods.utils.inv_probit with its 1e-3 jitter."""
    jitter = 1e-3
    return 0.5 * (1.0 + erf(x / np.sqrt(2.0))) * (1 - 2 * jitter) + jitter


class Bernoulli(_QuadratureLikelihood):
    """gpflow.likelihoods.Bernoulli(invlink=inv_probit); logp = log(where(y == 1, p, 1 - p))."""

    name = "bernoulli"

    def __init__(self, invlink=inv_probit):
        self.invlink = invlink

    def _logp(self, F, Y):
        p = inv_probit(F)
        return np.log(np.where(Y == 1, p, 1 - p))

    def _dlogp(self, F, Y):
        p = inv_probit(F)
        dp = (1 - 2e-3) * np.exp(-0.5 * np.square(F)) / np.sqrt(2 * np.pi)
        return np.where(Y == 1, dp / p, -dp / (1 - p))


class StudentT(_QuadratureLikelihood):
    """gpflow.likelihoods.StudentT(scale=1.0, df=3.0); logdensities.student_t."""

    name = "student_t"

    def __init__(self, scale=1.0, df=3.0):
        self.scale = _param(scale)
        self.df = float(df)

    def _logp(self, F, Y):
        df, sc = self.df, float(self.scale)
        const = gammaln((df + 1.0) * 0.5) - gammaln(df * 0.5) - 0.5 * (np.log(np.square(sc)) + np.log(df) + np.log(np.pi))
        return const - 0.5 * (df + 1.0) * np.log(1.0 + (1.0 / df) * np.square((Y - F) / sc))

    def _dlogp(self, F, Y):
        df, sc = self.df, float(self.scale)
        r = Y - F
        return (df + 1.0) * r / (df * sc * sc + np.square(r))


class Softmax:
    """gpflow.likelihoods.Softmax(num_classes) [GPflow-recalled]: a MonteCarloLikelihood with num_monte_carlo_points = 100.
    variational_expectations = mean over S draws of log softmax(Fmu + sqrt(Fvar) * eps)[y], eps ~ N(0, I) of shape [S, N, L]
    (MonteCarloLikelihood._mc_quadrature); Y holds integer class labels [N, 1] (sparse_softmax_cross_entropy_with_logits).
    GPflow draws eps afresh in every call (tf.random.normal); for a reproducible comparison the draws are an attribute here
    (`epsilon`, GPflow's own optional argument), and the gradients the reference takes by tf.GradientTape (tsvgp.py:256-259)
    are the analytic derivatives of the same Monte-Carlo sum for fixed eps."""

    name = "softmax"

    def __init__(self, num_classes, num_monte_carlo_points=100, epsilon=None):
        self.num_classes = int(num_classes)
        self.num_monte_carlo_points = int(num_monte_carlo_points)
        self.epsilon = epsilon

    def _eps(self, N):
        if self.epsilon is None or self.epsilon.shape != (self.num_monte_carlo_points, N, self.num_classes):
            raise ValueError("Softmax oracle: set .epsilon to an array [S, N, L] (GPflow would draw tf.random.normal here)")
        return self.epsilon

    def variational_expectations(self, Fmu, Fvar, Y):
        return self.ve_and_grads(Fmu, Fvar, Y)[0]

    def ve_and_grads(self, Fmu, Fvar, Y):
        N, L = Fmu.shape
        eps = self._eps(N)
        sd = np.sqrt(Fvar)
        F = Fmu[None] + sd[None] * eps                         # [S, N, L]
        F = F - F.max(axis=-1, keepdims=True)
        logp = F - np.log(np.sum(np.exp(F), axis=-1, keepdims=True))
        y = np.asarray(Y[:, 0], dtype=int)
        onehot = np.eye(L)[y]                                  # [N, L]
        ve = np.mean(np.take_along_axis(logp, y[None, :, None], axis=-1)[..., 0], axis=0)
        d = onehot[None] - np.exp(logp)                        # d log softmax[y] / d f
        g_mean = np.mean(d, axis=0)
        g_var = np.mean(d * eps, axis=0) / (2.0 * sd)
        return ve, g_mean, g_var


# ----------------------------------------------------------------------------------------
# TF linear-algebra ops, as the reference calls them.
# ----------------------------------------------------------------------------------------
def cholesky(A):
    """tf.linalg.cholesky (lower; raises on failure)."""
    return sla.cholesky(A, lower=True, check_finite=False)


def triangular_solve(L, B, lower=True, adjoint=False):
    """tf.linalg.triangular_solve."""
    return sla.solve_triangular(L, B, lower=lower, trans=1 if adjoint else 0, check_finite=False)


def cholesky_solve(chol, rhs):
    """tf.linalg.cholesky_solve: forward then back substitution."""
    return triangular_solve(chol, triangular_solve(chol, rhs), adjoint=True)


# ----------------------------------------------------------------------------------------
# Reference functions on the path.
# ----------------------------------------------------------------------------------------
def posterior_from_dense_site(K, lambda_1, lambda_2_sqrt):
    """reference src/util.py:349-391 (L = 1 .. any; loops the leading latent axis)."""
    Lt = np.asarray(lambda_2_sqrt)
    M = K.shape[-1]
    assert Lt.ndim == 3 and Lt.shape[1:] == (M, M) and lambda_1.shape == (M, Lt.shape[0])
    Id = np.eye(M)
    C = cholesky(K)  # :377
    m_q = np.empty_like(lambda_1)
    chol_S = np.empty_like(Lt)
    for l in range(Lt.shape[0]):
        L = Lt[l]
        CtL = C.T @ L  # :380
        W = Id + CtL.T @ CtL  # :381
        chol_W = cholesky(W)  # :382
        LtK = L.T @ K  # :385
        iwLtK = triangular_solve(chol_W, LtK)  # :386
        S_q = K - iwLtK.T @ iwLtK  # :387
        chol_S[l] = cholesky(S_q)  # :388
        m_q[:, l] = S_q @ lambda_1[:, l]  # :389
    return m_q, chol_S


def base_conditional(Kmn, Kmm, Knn, f, q_sqrt):
    """gpflow.conditionals.util.base_conditional, full_cov=False, white=False. Returns [N,L],[N,L]."""
    Lm = cholesky(Kmm)
    A = triangular_solve(Lm, Kmn)  # [M,N]
    fvar = Knn - np.sum(np.square(A), axis=0)  # [N]
    A = triangular_solve(Lm, A, adjoint=True)  # unwhitened
    fmean = A.T @ f  # [N,L]
    Lq = np.tril(q_sqrt)  # band_part(q_sqrt, -1, 0)
    fvar_out = np.empty_like(fmean)
    for l in range(Lq.shape[0]):
        LTA = Lq[l].T @ A  # [M,N]
        fvar_out[:, l] = fvar + np.sum(np.square(LTA), axis=0)
    return fmean, fvar_out


def gauss_kl(q_mu, q_sqrt, K):
    """gpflow.kullback_leiblers.gauss_kl, non-diagonal q_sqrt [L,M,M], shared K [M,M]."""
    M, L = q_mu.shape
    Lp = cholesky(K)
    alpha = triangular_solve(Lp, q_mu)
    Lq = np.tril(q_sqrt)
    mahalanobis = np.sum(np.square(alpha))
    constant = -float(M * L)
    logdet_qcov = np.sum(np.log(np.square(np.diagonal(Lq, axis1=-2, axis2=-1))))
    trace = 0.0
    for l in range(L):
        LpiLq = triangular_solve(Lp, Lq[l])
        trace += np.sum(np.square(LpiLq))
    twoKL = mahalanobis + constant - logdet_qcov + trace
    twoKL += L * np.sum(np.log(np.square(np.diagonal(Lp))))
    return 0.5 * twoKL


def gradient_transformation_mean_var_to_expectation(inputs, grads):
    """reference src/util.py:429-438."""
    return grads[0] - 2.0 * np.einsum("lmo,ol->ml", grads[1], inputs), grads[1]


def conditional_from_precision_sites(Kuu_, Kff, Kuf_, l, L):
    """reference src/util.py:91-185 (the `new_predict_f` algebra); L given, one latent per leading index."""
    M = Kuu_.shape[-1]
    Id = np.eye(M)
    C = cholesky(Kuu_)
    means, covs = [], []
    for i in range(L.shape[0]):
        CtL = C.T @ L[i]
        W = Id + CtL.T @ CtL
        chol_W = cholesky(W)
        D = triangular_solve(chol_W, L[i].T)
        tmp = D @ Kuf_
        mean = Kuf_.T @ l[:, i] - np.sum((D @ (Kuu_ @ l[:, i]))[:, None] * tmp, axis=0)
        cov = Kff[:, 0] - np.sum(np.square(tmp), axis=0)
        means.append(mean)
        covs.append(cov)
    return np.stack(means, -1), np.stack(covs, -1)


class OracleTSVGP:
    """The reference's ``t_SVGP`` (tsvgp.py:117-304) on NumPy arrays, same public surface."""

    def __init__(self, kernel, likelihood, inducing_variable, *, mean_function=None, num_latent_gps=1,
                 lambda_1=None, lambda_2_sqrt=None, num_data=None, force=False):
        self.kernel = kernel
        self.likelihood = likelihood
        self.mean_function = mean_function
        self.num_latent_gps = num_latent_gps
        self.num_data = num_data
        self.inducing_variable = inducing_variable if hasattr(inducing_variable, "Z") else InducingPoints(inducing_variable)
        self.num_inducing = self.inducing_variable.num_inducing
        M = self.num_inducing
        if lambda_1 is None:  # tsvgp.py:174
            lambda_1 = np.zeros((M, self.num_latent_gps))
        if lambda_2_sqrt is None:  # tsvgp.py:175-180
            lambda_2_sqrt = np.array([-np.eye(M) * 1e-10 for _ in range(self.num_latent_gps)])
        else:
            assert lambda_2_sqrt.ndim == 3  # tsvgp.py:182
            self.num_latent_gps = lambda_2_sqrt.shape[0]
        self.lambda_1 = np.array(lambda_1, dtype=np.float64)
        # sites.py:63 — triangular() transform keeps the lower triangle only
        self.lambda_2_sqrt = np.tril(np.array(lambda_2_sqrt, dtype=np.float64))
        self.whiten = False
        self.force = force

    @property
    def lambda_2(self):  # tsvgp.py:197-200
        return self.lambda_2_sqrt @ np.swapaxes(self.lambda_2_sqrt, -1, -2)

    def _mean_fn(self, X):
        if self.mean_function is None:
            return 0.0
        return np.asarray(self.mean_function(X), dtype=np.float64).reshape(X.shape[0], -1)

    def get_mean_chol_cov_inducing_posterior(self):  # tsvgp.py:202-212
        K_uu = Kuu(self.inducing_variable, self.kernel, jitter=DEFAULT_JITTER)
        return posterior_from_dense_site(K_uu, self.lambda_1, self.lambda_2_sqrt)

    def prior_kl(self):  # tsvgp.py:65-70
        q_mu, q_sqrt = self.get_mean_chol_cov_inducing_posterior()
        K = Kuu(self.inducing_variable, self.kernel, jitter=DEFAULT_JITTER)
        return gauss_kl(q_mu, q_sqrt, K)

    def predict_f(self, Xnew, full_cov=False, full_output_cov=False):  # tsvgp.py:97-114
        assert not full_cov and not full_output_cov
        Xnew = np.asarray(Xnew, dtype=np.float64)
        q_mu, q_sqrt = self.get_mean_chol_cov_inducing_posterior()
        Kmm = Kuu(self.inducing_variable, self.kernel, jitter=DEFAULT_JITTER)
        Kmn = Kuf(self.inducing_variable, self.kernel, Xnew)
        Knn = self.kernel.K_diag(Xnew)
        mu, var = base_conditional(Kmn, Kmm, Knn, q_mu, q_sqrt)
        if not np.all(var > 0):  # tf.debugging.assert_positive(var), :113
            raise FloatingPointError("predict_f: non-positive predictive variance")
        return mu + self._mean_fn(Xnew), var

    def new_predict_f(self, Xnew):  # tsvgp.py:215-232
        Xnew = np.asarray(Xnew, dtype=np.float64)
        K_uu = Kuu(self.inducing_variable, self.kernel, jitter=DEFAULT_JITTER)
        K_uf = Kuf(self.inducing_variable, self.kernel, Xnew)
        K_ff = self.kernel.K_diag(Xnew)[..., None]
        mu, var = conditional_from_precision_sites(K_uu, K_ff, K_uf, self.lambda_1, self.lambda_2_sqrt)
        if not np.all(var > 0):
            raise FloatingPointError("new_predict_f: non-positive predictive variance")
        return mu + self._mean_fn(Xnew), var

    def elbo(self, data):  # tsvgp.py:79-95
        X, Y = data
        X = np.asarray(X, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64)
        kl = self.prior_kl()
        f_mean, f_var = self.predict_f(X)
        var_exp = self.likelihood.variational_expectations(f_mean, f_var, Y)
        scale_ = self.num_data / X.shape[0] if self.num_data is not None else 1.0
        return np.sum(var_exp) * scale_ - kl

    def natgrad_step(self, data, lr=0.1, jitter=1e-9):  # tsvgp.py:234-304
        X, Y = data
        X = np.asarray(X, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64)
        mean, var = self.predict_f(X)  # :246
        meanZ, _ = self.predict_f(np.asarray(self.inducing_variable.Z))  # :254
        _, g_mean, g_var = self.likelihood.ve_and_grads(mean, var, Y)  # :256-259
        g_var = np.minimum(g_var, -1e-8)  # :262-263
        M = self.num_inducing
        Id = np.eye(M)
        K_uu = Kuu(self.inducing_variable, self.kernel)  # :268 (no jitter)
        K_uf = Kuf(self.inducing_variable, self.kernel, X)  # :269
        chol_Kuu = cholesky(K_uu + Id * jitter)  # :270
        A = cholesky_solve(chol_Kuu, K_uf).T  # :271  [N,M]
        G1 = A.T @ g_mean  # :279  einsum('nml,nl->ml') with A tiled over l
        G2 = np.stack([(A * g_var[:, l:l + 1]).T @ A for l in range(g_var.shape[1])])  # :280
        grad_mu = gradient_transformation_mean_var_to_expectation(meanZ, (G1, G2))  # :284
        scale_ = self.num_data / X.shape[0] if self.num_data is not None else 1.0  # :286-291
        lambda_2 = -0.5 * self.lambda_2  # :293
        lambda_1 = (1 - lr) * self.lambda_1 + lr * scale_ * grad_mu[0]  # :296
        lambda_2 = (1 - lr) * lambda_2 + lr * scale_ * grad_mu[1]  # :297
        new_sqrt = np.stack([-cholesky(-2.0 * lambda_2[l] + Id * jitter) for l in range(lambda_2.shape[0])])  # :300
        self.lambda_1 = lambda_1  # :302
        self.lambda_2_sqrt = new_sqrt  # :303


# ----------------------------------------------------------------------------------------
# Closed-form exact GP regression (what the reference's tests compare against through
# gpflow.models.GPR; used to pin the oracle by the reference's own properties).
# ----------------------------------------------------------------------------------------
def gpr_log_marginal_likelihood(kernel, X, Y, noise_variance):
    """gpflow.models.GPR.log_marginal_likelihood (zero mean function)."""
    N = X.shape[0]
    K = kernel.K(X) + noise_variance * np.eye(N)
    L = cholesky(K)
    alpha = triangular_solve(L, Y)
    return float(-0.5 * np.sum(np.square(alpha)) - np.sum(np.log(np.diag(L))) - 0.5 * N * np.log(2 * np.pi))


def gpr_predict_f(kernel, X, Y, noise_variance, Xnew):
    """gpflow.models.GPR.predict_f, full_cov=False."""
    N = X.shape[0]
    Kmm = kernel.K(X) + noise_variance * np.eye(N)
    Kmn = kernel.K(X, Xnew)
    Lm = cholesky(Kmm)
    A = triangular_solve(Lm, Kmn)
    fvar = kernel.K_diag(Xnew) - np.sum(np.square(A), axis=0)
    fmean = A.T @ triangular_solve(Lm, Y)
    return fmean, fvar[:, None]


def predict_y(model, Xnew):
    """gpflow.models.GPModel.predict_y -> likelihood.predict_mean_and_var [GPflow-recalled]: Gaussian (mu, var + s2);
    Bernoulli(inv_probit): p = inv_probit(mu / sqrt(1 + var)), p - p^2; StudentT by quadrature = (mu, var + scale^2 df/(df-2))."""
    mu, var = model.predict_f(Xnew)
    lik = model.likelihood
    if isinstance(lik, Gaussian):
        return mu, var + float(lik.variance)
    if isinstance(lik, Bernoulli):
        p = inv_probit(mu / np.sqrt(1.0 + var))
        return p, p - np.square(p)
    z, w = gh_points_and_weights(lik.n_gh)   # ScalarLikelihood._predict_mean_and_var: quadrature of E[y|f] = f and Var[y|f] + f^2
    F = mu[..., None] + np.sqrt(var)[..., None] * z
    cv = float(lik.scale) ** 2 * lik.df / (lik.df - 2.0)
    Ey = np.sum(F * w, axis=-1)
    Ey2 = np.sum((cv + np.square(F)) * w, axis=-1)
    return Ey, Ey2 - np.square(Ey)


def predict_log_density(model, data):
    """gpflow.models.GPModel.predict_log_density [GPflow-recalled]."""
    from scipy.special import logsumexp
    X, Y = data
    mu, var = model.predict_f(X)
    lik = model.likelihood
    if isinstance(lik, Gaussian):
        v = var + float(lik.variance)
        return np.sum(-0.5 * (np.log(2 * np.pi) + np.log(v) + np.square(Y - mu) / v), axis=-1)
    if isinstance(lik, Bernoulli):
        p, _ = predict_y(model, X)
        return np.sum(np.log(np.where(Y == 1, p, 1 - p)), axis=-1)
    z, w = gh_points_and_weights(lik.n_gh)
    F = mu[..., None] + np.sqrt(var)[..., None] * z
    return np.sum(logsumexp(lik._logp(F, Y[..., None]) + np.log(w), axis=-1), axis=-1)


# ----------------------------------------------------------------------------------------
# M-step: gradients of the ELBO w.r.t. kernel variance, lengthscales, inducing inputs and
# likelihood parameters with the sites held fixed.  The reference obtains them by TensorFlow
# autodiff through `elbo` (tsvgp.py:79-95; callers docs/notebooks/mnist.py:161-163,188-189;
# pinned by tests/models/test_tsvgp.py:168-188).  Here: the analytic derivative of the same
# function (DESIGN.md "M-step gradients"), validated against central differences of
# OracleTSVGP.elbo in tests/test_elbo_gradients_cpu.py.
# ----------------------------------------------------------------------------------------
def _kernel_dr2(kernel, r2):
    """dK/d(r^2) of the stationary kernels (K as a function of the scaled squared distance)."""
    var = float(kernel.variance)
    if isinstance(kernel, SquaredExponential):
        return -0.5 * var * np.exp(-0.5 * r2)
    r = np.sqrt(np.maximum(r2, 1e-36))
    sqrt5 = np.sqrt(5.0)
    return -(5.0 / 6.0) * var * (1.0 + sqrt5 * r) * np.exp(-sqrt5 * r)


def elbo_gradients(model, data):
    """-> (elbo, dict(variance=, lengthscales=[D], Z=[M, D], likelihood=...)) for num_latent_gps = 1, Zero mean function.
    `likelihood` is d/d(variance) for Gaussian, d/d(scale) for StudentT, None for Bernoulli."""
    X, Y = (np.asarray(a, dtype=np.float64) for a in data)
    kernel, lik = model.kernel, model.likelihood
    Z = np.asarray(model.inducing_variable.Z)
    M, D = Z.shape
    N = X.shape[0]
    ls = np.broadcast_to(np.asarray(kernel.lengthscales, dtype=np.float64).reshape(-1), (D,)).copy()
    var = float(kernel.variance)
    s = model.num_data / N if model.num_data is not None else 1.0
    lam1, L2 = model.lambda_1[:, 0], model.lambda_2_sqrt[0]
    Lam2 = L2 @ L2.T
    K = kernel.K(Z)
    K6 = K + DEFAULT_JITTER * np.eye(M)
    Kuf_ = kernel.K(Z, X)
    IK = np.eye(M) + Lam2 @ K6
    alpha = np.linalg.solve(IK, lam1)                 # K6^-1 m_q
    Q = np.linalg.solve(IK, Lam2)                     # (Lam2^-1 + K6)^-1, symmetric
    Q = 0.5 * (Q + Q.T)
    m = K6 @ alpha
    mu = Kuf_.T @ alpha
    v = var - np.sum(Kuf_ * (Q @ Kuf_), axis=0)
    ve, g, h = lik.ve_and_grads(mu[:, None], v[:, None], Y)
    g, h = g[:, 0], h[:, 0]
    kl = 0.5 * (m @ alpha - np.trace(Q @ K6) + np.linalg.slogdet(IK)[1])
    elbo = s * np.sum(ve) - kl
    b = Kuf_ @ g
    B = (Kuf_ * h) @ Kuf_.T
    G_uf = s * (np.outer(alpha, g) - 2.0 * (Q @ Kuf_) * h)                       # dELBO/dKuf
    sym = lambda A: 0.5 * (A + A.T)  # noqa: E731
    G_K = s * (Q @ B @ Q - sym(np.outer(Q @ b, alpha))) - 0.5 * (np.outer(alpha, alpha) - 2.0 * sym(np.outer(Q @ m, alpha)) + Q @ K6 @ Q)
    Zs, Xs = Z / ls, X / ls
    r2_uf = square_distance(Zs, Xs)
    r2_uu = square_distance(Zs, None)
    E_uf = G_uf * _kernel_dr2(kernel, r2_uf)
    E_uu = G_K * _kernel_dr2(kernel, r2_uu)
    np.fill_diagonal(E_uu, 0.0)                       # r = 0 on the diagonal: no dependence on Z or the lengthscales
    d_ls = np.empty(D)
    d_Z = np.empty((M, D))
    for d in range(D):
        duf = Zs[:, d:d + 1] - Xs[None, :, d]
        duu = Zs[:, d:d + 1] - Zs[None, :, d]
        d_ls[d] = -(2.0 / ls[d]) * (np.sum(E_uf * duf * duf) + np.sum(E_uu * duu * duu))
        d_Z[:, d] = (2.0 / ls[d]) * (np.sum(E_uf * duf, axis=1) + 2.0 * np.sum(E_uu * duu, axis=1))
    d_var = (np.sum(G_uf * Kuf_) + np.sum(G_K * K)) / var + s * np.sum(h)
    if isinstance(lik, Gaussian):
        s2 = float(lik.variance)
        d_lik = s * np.sum(-0.5 / s2 + 0.5 * (np.square(Y[:, 0] - mu) + v) / (s2 * s2))
    elif isinstance(lik, StudentT):
        z, w = gh_points_and_weights(lik.n_gh)
        F = mu[:, None] + np.sqrt(v)[:, None] * z
        r_ = Y - F
        sc, df = float(lik.scale), lik.df
        d_lik = s * np.sum((-1.0 / sc + (df + 1.0) * np.square(r_) / (sc * (df * sc * sc + np.square(r_)))) * w)
    else:
        d_lik = None
    if np.asarray(kernel.lengthscales).size == 1:
        d_ls = np.array([np.sum(d_ls)])
    return elbo, dict(variance=d_var, lengthscales=d_ls, Z=d_Z, likelihood=d_lik)


# ----------------------------------------------------------------------------------------
# Whitened sibling: t_SVGP_white (reference src/models/tsvgp_white.py) and its three helpers
# (reference src/util.py:11-88, 239-291, 394-426).  Second "next" row of SURVEY 8f.
# ----------------------------------------------------------------------------------------
def conditional_from_precision_sites_white(Kuu_, Kff, Kuf_, l, L2, jitter=1e-9):
    """reference src/util.py:11-88 (L2 given): mean [N, L], cov [N, L]."""
    M = Kuu_.shape[-1]
    Id = np.eye(M)
    LA = cholesky(Kuu_)
    tmp2 = triangular_solve(LA, Kuf_)
    means, covs = [], []
    for i in range(L2.shape[0]):
        LR = cholesky(L2[i] + Kuu_ + Id * jitter)
        tmp1 = triangular_solve(LR, Kuf_)
        covs.append(Kff[:, 0] - (np.sum(np.square(tmp2), axis=0) - np.sum(np.square(tmp1), axis=0)))
        means.append(Kuf_.T @ cholesky_solve(LR, l[:, i]))
    return np.stack(means, -1), np.stack(covs, -1)


def kl_from_precision_sites_white(A, l, L2):
    """reference src/util.py:239-291 (L2 given)."""
    M = A.shape[-1]
    LA = cholesky(A)
    kl = 0.0
    for i in range(L2.shape[0]):
        LR = cholesky(L2[i] + A)
        log_det = np.sum(np.log(np.square(np.diag(LR)))) - np.sum(np.log(np.square(np.diag(LA))))
        tmp = triangular_solve(LR, LA)
        trace_plus_const = np.sum(np.square(tmp)) - M
        mahalanobis = np.sum(np.square(LA.T @ cholesky_solve(LR, l[:, i])))
        kl += 0.5 * (log_det + trace_plus_const + mahalanobis)
    return kl


def posterior_from_dense_site_white(K, lambda_1, lambda_2, jitter=1e-9):
    """reference src/util.py:394-426."""
    M = K.shape[-1]
    Id = np.eye(M)
    m_q = np.empty_like(lambda_1)
    chol_S = np.empty_like(lambda_2)
    for i in range(lambda_2.shape[0]):
        LR = cholesky(K + lambda_2[i] + Id * jitter)
        iLRK = triangular_solve(LR, K)
        chol_S[i] = cholesky(iLRK.T @ iLRK)
        m_q[:, i] = K @ cholesky_solve(LR, lambda_1[:, i])
    return m_q, chol_S


class OracleTSVGPWhite:
    """The reference's ``t_SVGP_white`` (src/models/tsvgp_white.py:23-246) on NumPy arrays."""

    def __init__(self, kernel, likelihood, inducing_variable, *, mean_function=None, num_latent_gps=1, lambda_1=None,
                 lambda_2=None, num_data=None):
        self.kernel, self.likelihood, self.mean_function, self.num_data = kernel, likelihood, mean_function, num_data
        self.inducing_variable = inducing_variable if hasattr(inducing_variable, "Z") else InducingPoints(inducing_variable)
        self.num_inducing = M = self.inducing_variable.num_inducing
        self.num_latent_gps = num_latent_gps
        if lambda_2 is None:                                                  # :79-85
            lambda_2 = np.array([np.eye(M) * 1e-10 for _ in range(num_latent_gps)])
        else:
            assert lambda_2.ndim == 3
            self.num_latent_gps = lambda_2.shape[0]
        self.lambda_1 = np.zeros((M, self.num_latent_gps)) if lambda_1 is None else np.array(lambda_1, dtype=np.float64)
        self.lambda_2 = np.array(lambda_2, dtype=np.float64)

    def _mean_fn(self, X):
        return 0.0 if self.mean_function is None else np.asarray(self.mean_function(X), dtype=np.float64).reshape(X.shape[0], -1)

    def get_mean_chol_cov_inducing_posterior(self):                           # :99-109
        return posterior_from_dense_site_white(Kuu(self.inducing_variable, self.kernel, jitter=DEFAULT_JITTER), self.lambda_1, self.lambda_2)

    def prior_kl(self):                                                       # :116-120
        return kl_from_precision_sites_white(Kuu(self.inducing_variable, self.kernel, jitter=DEFAULT_JITTER), self.lambda_1, self.lambda_2)

    def predict_f(self, Xnew, full_cov=False, full_output_cov=False):         # :122-132
        Xnew = np.asarray(Xnew, dtype=np.float64)
        K_uu = Kuu(self.inducing_variable, self.kernel, jitter=DEFAULT_JITTER)
        K_uf = Kuf(self.inducing_variable, self.kernel, Xnew)
        K_ff = self.kernel.K_diag(Xnew)[..., None]
        mu, var = conditional_from_precision_sites_white(K_uu, K_ff, K_uf, self.lambda_1, self.lambda_2)
        if not np.all(var > 0):
            raise FloatingPointError("predict_f: non-positive predictive variance")
        return mu + self._mean_fn(Xnew), var

    def predict_f_extra_data(self, Xnew, extra_data, jitter=DEFAULT_JITTER):  # :134-158
        Xnew = np.asarray(Xnew, dtype=np.float64)
        grad_mu = self.compute_data_natural_params(extra_data)
        lambda_2 = -0.5 * self.lambda_2
        K_uu = Kuu(self.inducing_variable, self.kernel, jitter=jitter)
        lambda_1c = self.lambda_1 + K_uu @ grad_mu[0]
        lambda_2c = -2 * (lambda_2 + np.stack([K_uu @ grad_mu[1][l] @ K_uu for l in range(lambda_2.shape[0])]))
        K_uf = Kuf(self.inducing_variable, self.kernel, Xnew)
        K_ff = self.kernel.K_diag(Xnew)[..., None]
        mu, var = conditional_from_precision_sites_white(K_uu, K_ff, K_uf, lambda_1c, lambda_2c)
        return mu + self._mean_fn(Xnew), var

    def elbo(self, data):                                                     # :162-177
        X, Y = (np.asarray(a, dtype=np.float64) for a in data)
        kl = self.prior_kl()
        f_mean, f_var = self.predict_f(X)
        var_exp = self.likelihood.variational_expectations(f_mean, f_var, Y)
        scale_ = self.num_data / X.shape[0] if self.num_data is not None else 1.0
        return np.sum(var_exp) * scale_ - kl

    def compute_data_natural_params(self, data, jitter=1e-9):                 # :183-212 (no clipping of the variance gradient here)
        X, Y = (np.asarray(a, dtype=np.float64) for a in data)
        mean, var = self.predict_f(X)
        meanZ, _ = self.predict_f(np.asarray(self.inducing_variable.Z))
        _, g_mean, g_var = self.likelihood.ve_and_grads(mean, var, Y)
        M = self.num_inducing
        K_uu = Kuu(self.inducing_variable, self.kernel)
        K_uf = Kuf(self.inducing_variable, self.kernel, X)
        chol_Kuu = cholesky(K_uu + np.eye(M) * jitter)
        A = cholesky_solve(chol_Kuu, K_uf).T
        G1 = A.T @ g_mean
        G2 = np.stack([(A * g_var[:, l:l + 1]).T @ A for l in range(g_var.shape[1])])
        return gradient_transformation_mean_var_to_expectation(meanZ, (G1, G2))

    def natgrad_step(self, dataset, lr=0.1, jitter=1e-9):                     # :215-246
        X, Y = dataset
        grad_mu = self.compute_data_natural_params((X, Y))
        K_uu = Kuu(self.inducing_variable, self.kernel)
        scale_ = self.num_data / np.asarray(X).shape[0] if self.num_data is not None else 1.0
        lambda_2 = -0.5 * self.lambda_2
        self.lambda_1 = (1.0 - lr) * self.lambda_1 + lr * scale_ * K_uu @ grad_mu[0]
        lambda_2 = (1.0 - lr) * lambda_2 + lr * scale_ * np.stack([K_uu @ grad_mu[1][l] @ K_uu for l in range(lambda_2.shape[0])])
        self.lambda_2 = -2.0 * lambda_2
