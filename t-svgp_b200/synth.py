"""
Synthetic inputs of the five BASELINE.json configs: shapes, seeds, kernels, lengthscales chosen so that
cond(Kuu) stays in 1e2..1e4, likelihoods, learning rates.  Pure NumPy; shared by tests/ and bench.py.
"""
import numpy as np

CONFIGS = {
    # name: N (dataset), Nb (minibatch), M, D, kernel, lengthscale, likelihood, lik params, lr, seed
    "cfg1": dict(N=10_000, Nb=10_000, M=50, D=1, kernel="SquaredExponential", ls=0.05, lik="Gaussian", lik_args=dict(variance=0.1), lr=1.0, seed=1001),
    "cfg2": dict(N=100_000, Nb=10_000, M=500, D=8, kernel="SquaredExponential", ls=1.414, lik="Bernoulli", lik_args={}, lr=0.5, seed=1002),
    "cfg3": dict(N=10_000_000, Nb=1_000_000, M=2048, D=16, kernel="Matern52", ls=3.0, lik="Gaussian", lik_args=dict(variance=0.1), lr=0.5, seed=1003),
    "cfg4": dict(N=1_000_000, Nb=1_000_000, M=8192, D=8, kernel="SquaredExponential", ls=0.8, lik="Gaussian", lik_args=dict(variance=0.1), lr=0.5, seed=1004),
    "cfg5": dict(N=50_000_000, Nb=2_000_000, M=4096, D=32, kernel="SquaredExponential", ls=2.83, lik="StudentT", lik_args=dict(scale=0.3, df=3.0), lr=0.3, seed=1005),
}


def describe(name, **over):
    c = dict(CONFIGS[name])
    c.update({k: v for k, v in over.items() if v is not None})
    c["name"] = name
    return c


def make_minibatch(cfg, n_rows=None, M=None, seed_offset=0):
    """One minibatch (X [n, D], Y [n, 1]) and the inducing inputs Z [M, D] of a config; `n_rows` / `M` shrink it."""
    n = int(n_rows if n_rows is not None else cfg["Nb"])
    M = int(M if M is not None else cfg["M"])
    D = cfg["D"]
    rng = np.random.default_rng(cfg["seed"] + seed_offset)
    if cfg["name"] == "cfg1":
        X = rng.uniform(-1.0, 1.0, size=(n, 1))
        Z = np.linspace(-1.0, 1.0, M)[:, None]
        Y = np.sin(15.0 * X) + np.sqrt(0.1) * rng.standard_normal((n, 1))
        return X, Y, Z
    X = rng.standard_normal((max(n, M), D))
    Z = X[:M].copy()  # Z = X[:M] as the reference's scripts do (experiments/uci_regression.py:208)
    X = X[:n]
    eps = rng.standard_normal((n, 1))
    if cfg["lik"] == "Bernoulli":
        Y = (np.sin(X.sum(1, keepdims=True)) + 0.3 * eps > 0).astype(np.float64)
    elif cfg["lik"] == "StudentT":
        t = rng.standard_t(3.0, size=(n, 1))
        Y = np.sin(X.sum(1, keepdims=True) / np.sqrt(D)) + 0.3 * t
    else:
        Y = np.sin(X.sum(1, keepdims=True) / np.sqrt(D)) + np.sqrt(0.1) * eps
    return np.ascontiguousarray(X), np.ascontiguousarray(Y), Z


def build_objects(cfg, ns):
    """Kernel and likelihood objects from a namespace that has GPflow's class names (gpflow itself, the oracle module, or
    tsvgp_b200.standins)."""
    kernel = getattr(ns, cfg["kernel"])(variance=1.0, lengthscales=cfg["ls"])
    lik = getattr(ns, cfg["lik"])(**cfg["lik_args"])
    return kernel, lik


def flops_per_point(M, D, Q=0, L=1):
    """SURVEY §8d algorithmic work per data point of one natgrad_step."""
    return L * 2 * M * M + M * (2 * D + 6) + L * 4 * M + L * Q * 60


def dense_flops(M):
    return 8.0 * M ** 3
