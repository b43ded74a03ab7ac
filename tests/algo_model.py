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


# ---- the three statistics routes ------------------------------------------------------------------------------
def natural_gradients(K, Kuf, g, h, jitter=1e-9, route="fused"):
    """(G1, G2) of tsvgp.py:271-281 as each device route forms them.
    fused    : B = Kuf diag(h) Kfu, then K9^-1 B K9^-1                      (2 M^2 flops per point, rounding ~ eps cond^2)
    whitened : t = C9^-1 k per slab, Bw = sum h t t^T, then C9^-T Bw C9^-1   (3 M^2, rounding ~ eps cond)
    exact    : a = C9^-T C9^-1 k per slab, G2 = sum h a a^T                  (4 M^2; a Gram product, as the reference)"""
    M = K.shape[0]
    C9 = np.linalg.cholesky(K + jitter * np.eye(M))
    C9inv = sla.solve_triangular(C9, np.eye(M), lower=True)
    if route == "fused":
        K9inv = C9inv.T @ C9inv
        return K9inv @ (Kuf @ g), K9inv @ ((Kuf * h) @ Kuf.T) @ K9inv
    Wt = C9inv @ Kuf
    if route == "whitened":
        return C9inv.T @ (Wt @ g), C9inv.T @ ((Wt * h) @ Wt.T) @ C9inv
    A = C9inv.T @ Wt
    return A @ g, (A * h) @ A.T


# ---- block schedule of diag_potrf_inv_blocked_kernel (csrc/kernels.cu) -------------------------------------------------------
def blocked_potrf_inv(A, q=32):
    """Cholesky L of a (4q x 4q) block and X = L^-1 in the kernel's order: per block step j the q x q diagonal sub-block is
    factored column by column with the elimination of L X = I riding on the same pivots; then panel L_ij = A_ij X_jj^T and the
    finished block row X_jc = X_jj Xcur_jc; then trailing A_ik -= L_ij L_kj^T and Xcur_ic -= L_ij X_jc (overwrite for c = j)."""
    nb = A.shape[0] // q
    As = {(i, j): A[i * q:(i + 1) * q, j * q:(j + 1) * q].copy() for i in range(nb) for j in range(i + 1)}
    Xs = {}
    rows = np.arange(q)
    for j in range(nb):
        a, x = np.tril(As[j, j]).copy(), np.eye(q)
        for k in range(q):
            rs = 1.0 / np.sqrt(a[k, k])
            col = np.where(rows > k, a[:, k] * rs, 0.0)
            a[:, k] = np.where(rows == k, a[k, k] * rs, col)
            x[k, :] *= rs
            a[:, k + 1:] -= np.outer(col, col[k + 1:])
            x -= np.outer(col, x[k, :])
        As[j, j], Xs[j, j] = np.tril(a), np.tril(x)
        for i in range(j + 1, nb):
            As[i, j] = As[i, j] @ Xs[j, j].T
        for c in range(j):
            Xs[j, c] = Xs[j, j] @ Xs[j, c]
        for i in range(j + 1, nb):
            for k in range(j + 1, i + 1):
                As[i, k] -= As[i, j] @ As[k, j].T
            for c in range(j + 1):
                Xs[i, c] = -As[i, j] @ Xs[j, c] if c == j else Xs[i, c] - As[i, j] @ Xs[j, c]
    n = nb * q
    L, X = np.zeros((n, n)), np.zeros((n, n))
    for (i, j), blk in As.items():
        L[i * q:(i + 1) * q, j * q:(j + 1) * q] = blk
    for (i, j), blk in Xs.items():
        X[i * q:(i + 1) * q, j * q:(j + 1) * q] = blk
    return L, X
