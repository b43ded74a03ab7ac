"""
NumPy model of the algebra the CUDA path executes (NOT the oracle, NOT the product): used to
validate the reformulation against the reference-order oracle on the CPU and to debug device
intermediates.  See DESIGN.md "Algebra".

  K   = k(Z,Z);  K6 = K + 1e-6 I;  K9 = K + jitter I
  W   = I + L2^T K6 L2            (lower half formed directly; reference: util.py:380-381 via chol(K6))
  W   = Uw Uw^T                   (reverse Cholesky, Uw upper)  ->  T = L2 Uw^-T  (lower triangular)
  Q   = T T^T = L2 W^-1 L2^T      so  v_n = k_nn - |T^T k_n|^2           (util.py:175-184 route)
  alpha = lambda_1 - Q K6 lambda_1 = K6^-1 m ;  mu_n = k_n^T alpha ;  m = K6 alpha ;  mZ = K alpha
  B   = sum_n h_n k_n k_n^T ; b = sum_n g_n k_n ;  G2 = K9^-1 B K9^-1 ; G1 = K9^-1 b
  KL  = 1/2 ( m^T alpha - tr(Q K6) + logdet W )
"""
import numpy as np
import scipy.linalg as sla


def rev_cholesky_upper(W):
    """W = U U^T with U upper triangular."""
    J = np.arange(W.shape[0])[::-1]
    Lr = sla.cholesky(W[np.ix_(J, J)], lower=True)
    return Lr[np.ix_(J, J)]


def prepare(K, lambda_1, L2, jitter6=1e-6):
    M = K.shape[0]
    K6 = K + jitter6 * np.eye(M)
    W = np.eye(M) + L2.T @ (K6 @ L2)
    Uw = rev_cholesky_upper(W)
    Uinv = sla.solve_triangular(Uw, np.eye(M), lower=False)  # upper
    T = L2 @ Uinv.T  # lower x lower
    u = K6 @ lambda_1
    alpha = lambda_1 - T @ (T.T @ u)
    return dict(K6=K6, W=W, Uw=Uw, T=T, alpha=alpha)


def marginals(Kuf, kdiag, T, alpha):
    mu = Kuf.T @ alpha
    s = T.T @ Kuf
    var = kdiag - np.sum(np.square(s), axis=0)
    return mu, var


def kl(K6, T, alpha, Uw):
    m = K6 @ alpha
    CT = K6 @ T
    trQK = np.sum(T * CT)
    logdetW = 2.0 * np.sum(np.log(np.diag(Uw)))
    return 0.5 * (float(m[:, 0] @ alpha[:, 0]) - trQK + logdetW)


def local_statistics(Kuf, g, h):
    """What one rank accumulates over its rows (fused route): B = Kuf diag(h) Kfu, b = Kuf g.  Summed over ranks by the
    one all-reduce of the step."""
    return (Kuf * h) @ Kuf.T, Kuf @ g


def update_from_statistics(K, B, b, alpha, lambda_1, L2, lr, scale, jitter=1e-9):
    """The replicated dense phase after the all-reduce."""
    M = K.shape[0]
    C9 = sla.cholesky(K + jitter * np.eye(M), lower=True)
    C9inv = sla.solve_triangular(C9, np.eye(M), lower=True)
    K9inv = C9inv.T @ C9inv
    G2 = K9inv @ B @ K9inv
    G1 = K9inv @ b
    mZ = K @ alpha[:, 0]
    g0 = G1 - 2.0 * G2 @ mZ
    l1 = (1 - lr) * lambda_1[:, 0] + lr * scale * g0
    P = (1 - lr) * (L2 @ L2.T) + lr * scale * (-2.0 * G2) + jitter * np.eye(M)
    return l1[:, None], -sla.cholesky(P, lower=True)


def natgrad(K, Kuf, g, h, alpha, lambda_1, L2, lr, scale, jitter=1e-9, whiten=False):
    M = K.shape[0]
    K9 = K + jitter * np.eye(M)
    C9 = sla.cholesky(K9, lower=True)
    C9inv = sla.solve_triangular(C9, np.eye(M), lower=True)
    if whiten:
        Tt = C9inv @ Kuf
        Bw = (Tt * h) @ Tt.T
        G2 = C9inv.T @ Bw @ C9inv
        G1 = C9inv.T @ (Tt @ g)
    else:
        B = (Kuf * h) @ Kuf.T
        b = Kuf @ g
        K9inv = C9inv.T @ C9inv
        G2 = K9inv @ B @ K9inv
        G1 = K9inv @ b
    mZ = K @ alpha[:, 0]
    g0 = G1 - 2.0 * G2 @ mZ
    l1 = (1 - lr) * lambda_1[:, 0] + lr * scale * g0
    P = (1 - lr) * (L2 @ L2.T) + lr * scale * (-2.0 * G2) + jitter * np.eye(M)
    return l1[:, None], -sla.cholesky(P, lower=True)
