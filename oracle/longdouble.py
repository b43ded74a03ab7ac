"""
Extended-precision arbiter for the oracle.  TEST INFRASTRUCTURE ONLY (same rules as tsvgp_oracle.py).

One `natgrad_step` + `predict_f` in the reference's own operation order (src/models/tsvgp.py:97-114, 234-304;
src/util.py:349-391, 429-438), Gaussian likelihood, on `np.longdouble` (x87 80-bit, eps 1.1e-19) with hand-written
unblocked Cholesky and substitutions (NumPy's LAPACK wrappers do not take long double).  O(M^3 + N M^2) Python-level
vector operations: for M <= ~64.  It answers one question the float64 oracle cannot: when the CUDA path and the oracle
differ at the 1e-10 level on an ill-conditioned problem, which of the two is closer to the exact arithmetic result of
the reference's formulas.
"""
import numpy as np

LD = np.longdouble


def chol(A):
    A = np.array(A, dtype=LD)
    n = A.shape[0]
    L = np.zeros_like(A)
    for j in range(n):
        d = A[j, j] - np.dot(L[j, :j], L[j, :j])
        if not d > 0:
            raise FloatingPointError("not positive definite")
        L[j, j] = np.sqrt(d)
        if j + 1 < n:
            L[j + 1:, j] = (A[j + 1:, j] - L[j + 1:, :j] @ L[j, :j]) / L[j, j]
    return L


def solve_lower(L, B):
    B = np.array(B, dtype=LD)
    X = np.zeros_like(B)
    for i in range(L.shape[0]):
        X[i] = (B[i] - L[i, :i] @ X[:i]) / L[i, i]
    return X


def solve_upper(U, B):
    B = np.array(B, dtype=LD)
    X = np.zeros_like(B)
    for i in range(U.shape[0] - 1, -1, -1):
        X[i] = (B[i] - U[i, i + 1:] @ X[i + 1:]) / U[i, i]
    return X


def se_kernel(X, X2, variance, lengthscale):
    Xs, X2s = np.asarray(X, dtype=LD) / LD(lengthscale), np.asarray(X2, dtype=LD) / LD(lengthscale)
    d = -2 * (Xs @ X2s.T) + np.sum(Xs * Xs, 1)[:, None] + np.sum(X2s * X2s, 1)[None, :]
    return LD(variance) * np.exp(-d / 2)


def posterior(K6, lam1, L2):
    M = K6.shape[0]
    Id = np.eye(M, dtype=LD)
    C = chol(K6)
    CtL = C.T @ L2
    cW = chol(Id + CtL.T @ CtL)
    iw = solve_lower(cW, L2.T @ K6)
    S = K6 - iw.T @ iw
    return S @ lam1, chol(S)


def predict_f(X, Z, variance, lengthscale, lam1, L2):
    M = Z.shape[0]
    K6 = se_kernel(Z, Z, variance, lengthscale) + LD(1e-6) * np.eye(M, dtype=LD)
    m, R = posterior(K6, lam1, L2)
    Kmn = se_kernel(Z, X, variance, lengthscale)
    Lm = chol(K6)
    A = solve_lower(Lm, Kmn)
    fvar = LD(variance) - np.sum(A * A, 0)
    A = solve_upper(Lm.T, A)
    fmean = A.T @ m
    LTA = R.T @ A
    return fmean, fvar + np.sum(LTA * LTA, 0)


def natgrad_step_gaussian(X, Y, Z, variance, lengthscale, noise, lam1, L2, lr, scale=1.0, jitter=1e-9):
    """-> (lambda_1 [M], lambda_2_sqrt [M, M]) in long double; lam1 [M], L2 [M, M] lower."""
    X, Z = np.asarray(X), np.asarray(Z)
    lam1, L2, y = np.asarray(lam1, dtype=LD), np.asarray(L2, dtype=LD), np.asarray(Y, dtype=LD).reshape(-1)
    M = Z.shape[0]
    Id = np.eye(M, dtype=LD)
    mean, _ = predict_f(X, Z, variance, lengthscale, lam1, L2)
    meanZ, _ = predict_f(Z, Z, variance, lengthscale, lam1, L2)
    g_mean = (y - mean) / LD(noise)
    g_var = np.full_like(mean, -0.5 / LD(noise))
    K = se_kernel(Z, Z, variance, lengthscale)
    Kuf = se_kernel(Z, X, variance, lengthscale)
    C9 = chol(K + LD(jitter) * Id)
    A = solve_upper(C9.T, solve_lower(C9, Kuf)).T
    G1 = A.T @ g_mean
    G2 = (A * g_var[:, None]).T @ A
    g0 = G1 - 2 * (G2 @ meanZ)
    lam2 = -(L2 @ L2.T) / 2
    new1 = (1 - LD(lr)) * lam1 + LD(lr) * LD(scale) * g0
    new2 = (1 - LD(lr)) * lam2 + LD(lr) * LD(scale) * G2
    return new1, -chol(-2 * new2 + LD(jitter) * Id)


# ----------------------------------------------------------------------------------------------------------------------------
# Non-conjugate likelihoods (the Gauss-Hermite path of tsvgp.py:256-263) in extended precision: the arbiter for the quadrature
# sums, their analytic (mean, variance) gradients and the -1e-8 clip.  The quadrature RULE is the reference's: GPflow's 20 nodes
# sqrt(2) x_k and weights w_k / sqrt(pi) from numpy.polynomial.hermite.hermgauss (float64 constants), used here as exact numbers.
# erf has no long-double implementation in NumPy/SciPy: mpmath at 80 bits, element by element (fine for N * 20 <= ~10^4 values).
# ----------------------------------------------------------------------------------------------------------------------------
def _erf_ld(x):
    import mpmath
    mpmath.mp.prec = 80
    f = np.frompyfunc(lambda v: LD(str(mpmath.erf(mpmath.mpf(str(v))))), 1, 1)   # decimal strings carry all 64 mantissa bits (21 digits)
    return f(np.asarray(x, dtype=LD)).astype(LD)


def gh_rule(n_gh=20):
    """The rule constants exactly as the float64 implementations hold them (oracle: tsvgp_oracle.gh_points_and_weights; CUDA:
    tsvgp_set_likelihood): float64(sqrt(2) x_k), float64(w_k / sqrt(pi)), widened to long double."""
    x, w = np.polynomial.hermite.hermgauss(n_gh)
    return np.asarray(x * np.sqrt(2.0), dtype=LD), np.asarray(w / np.sqrt(np.pi), dtype=LD)


def ve_grads(lik, mean, var, y):
    """d E_q[log p(y | f)] / d mean, d / d var (clipped at -1e-8, tsvgp.py:262-263) per point, long double.
    lik = ("gaussian", variance) | ("bernoulli",) | ("student_t", scale, df)."""
    kind = lik[0]
    if kind == "gaussian":
        return (y - mean) / LD(lik[1]), np.minimum(np.full_like(mean, -0.5 / LD(lik[1])), LD(-1e-8))
    z, w = gh_rule(20)
    sd = np.sqrt(var)
    F = mean[:, None] + sd[:, None] * z[None, :]
    if kind == "bernoulli":            # inv_probit with GPflow's 1e-3 jitter; labels other than 1 count as class 0
        PI = LD("3.14159265358979323846264338327950288")
        p = LD(0.5) * (1 + _erf_ld(F / np.sqrt(LD(2)))) * (1 - 2 * LD(1e-3)) + LD(1e-3)
        dp = (1 - 2 * LD(1e-3)) * np.exp(-F * F / 2) / np.sqrt(2 * PI)
        d = np.where(y[:, None] == 1, dp / p, -dp / (1 - p))
    elif kind == "student_t":
        sc, df = LD(lik[1]), LD(lik[2])
        r = y[:, None] - F
        d = (df + 1) * r / (df * sc * sc + r * r)
    else:
        raise ValueError(kind)
    d = d * w[None, :]
    return np.sum(d, 1), np.minimum(np.sum(d * z[None, :], 1) / (2 * sd), LD(-1e-8))


def natgrad_step(X, Y, Z, variance, lengthscale, lik, lam1, L2, lr, scale=1.0, jitter=1e-9):
    """One natgrad_step (tsvgp.py:234-304) for any of the three likelihoods, SE kernel, long double; same order as
    natgrad_step_gaussian."""
    X, Z = np.asarray(X), np.asarray(Z)
    lam1, L2, y = np.asarray(lam1, dtype=LD), np.asarray(L2, dtype=LD), np.asarray(Y, dtype=LD).reshape(-1)
    M = Z.shape[0]
    Id = np.eye(M, dtype=LD)
    mean, var = predict_f(X, Z, variance, lengthscale, lam1, L2)
    meanZ, _ = predict_f(Z, Z, variance, lengthscale, lam1, L2)
    g_mean, g_var = ve_grads(lik, mean, var, y)
    K = se_kernel(Z, Z, variance, lengthscale)
    Kuf = se_kernel(Z, X, variance, lengthscale)
    C9 = chol(K + LD(jitter) * Id)
    A = solve_upper(C9.T, solve_lower(C9, Kuf)).T
    G1 = A.T @ g_mean
    G2 = (A * g_var[:, None]).T @ A
    g0 = G1 - 2 * (G2 @ meanZ)
    lam2 = -(L2 @ L2.T) / 2
    new1 = (1 - LD(lr)) * lam1 + LD(lr) * LD(scale) * g0
    new2 = (1 - LD(lr)) * lam2 + LD(lr) * LD(scale) * G2
    return new1, -chol(-2 * new2 + LD(jitter) * Id)
