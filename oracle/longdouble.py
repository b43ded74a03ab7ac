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
