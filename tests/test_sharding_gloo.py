"""
World-size-2 check (gloo, CPU) of the multi-GPU decomposition of SURVEY §8e: each rank accumulates the statistics of its
contiguous rows, ONE all-reduce sums them, the dense update is replicated — and the result equals the reference-order
full-minibatch step.  The arithmetic here is the NumPy model of the device algebra (tests/algo_model.py); the device
version of the same decomposition is exercised by tests/test_gpu_multi.py on >= 2 GPUs.
"""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out, split=False):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import tsvgp_oracle as orc
    from tests import algo_model as am
    import tsvgp_b200.synth as synth
    from tsvgp_b200 import shard_rows

    cfg = synth.describe("cfg2")
    N, M, num_data, lr = 1001, 48, 10_010, 0.5
    X, Y, Z = synth.make_minibatch(cfg, n_rows=N, M=M)
    kernel, lik = synth.build_objects(cfg, orc)
    ref = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(Z.copy()), num_data=num_data)
    ref.natgrad_step((X, Y), lr=lr)                       # non-trivial sites, same on every rank
    l1, L2 = ref.lambda_1.copy(), ref.lambda_2_sqrt[0].copy()

    K = kernel.K(Z)
    if split:
        # option "split_chains" (DESIGN 5): rank 0 alone builds the posterior factors and broadcasts them; it gets a smaller share
        # of the rows (tsvgp_b200.balance_weights turns measured phase times into shares: here a fixed 46 / 54)
        lo, hi = shard_rows(N, world, rank, [0.46, 0.54])
        pre = am.prepare(K, l1, L2) if rank == 0 else {"K6": K + 1e-6 * np.eye(M)}
        for key in ("T", "alpha", "Uw"):
            t = torch.from_numpy(np.ascontiguousarray(pre[key])) if rank == 0 else torch.zeros((M, 1) if key == "alpha" else (M, M), dtype=torch.float64)
            dist.broadcast(t, 0)
            pre[key] = t.numpy()
    else:
        lo, hi = shard_rows(N, world, rank)
        pre = am.prepare(K, l1, L2)
    Kuf = kernel.K(Z, X[lo:hi])
    mu, var = am.marginals(Kuf, kernel.K_diag(X[lo:hi]), pre["T"], pre["alpha"])
    ve, g, h = lik.ve_and_grads(mu, var[:, None], Y[lo:hi])
    h = np.minimum(h, -1e-8)
    B, b = am.local_statistics(Kuf, g[:, 0], h[:, 0])
    buf = torch.from_numpy(np.concatenate([B.ravel(), b, [ve.sum(), float(hi - lo)]]))
    dist.all_reduce(buf)                                   # the step's one collective
    buf = buf.numpy()
    B, b, ve_sum, n_global = buf[:M * M].reshape(M, M), buf[M * M:M * M + M], buf[-2], buf[-1]
    assert n_global == N
    scale = num_data / n_global                            # tsvgp.py:286-291 with the GLOBAL minibatch size
    n1, nL2 = am.update_from_statistics(K, B, b, pre["alpha"], l1, L2, lr, scale)
    elbo = scale * ve_sum - am.kl(pre["K6"], pre["T"], pre["alpha"], pre["Uw"])

    e_ref = ref.elbo((X, Y))
    ref.natgrad_step((X, Y), lr=lr)
    rel = lambda a, c: float(np.max(np.abs(a - c)) / np.max(np.abs(c)))  # noqa: E731
    out.put((rank, rel(n1, ref.lambda_1), rel(nL2 @ nL2.T, ref.lambda_2[0]), abs(elbo - e_ref) / abs(e_ref)))
    dist.destroy_process_group()


def test_two_rank_decomposition_matches_full_batch():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e1, e2, ee in res:
        assert e1 < 1e-9 and e2 < 1e-9 and ee < 1e-9, (rank, e1, e2, ee)


def test_two_rank_chain_split_with_uneven_shares_matches_full_batch():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out, True)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e1, e2, ee in res:
        assert e1 < 1e-9 and e2 < 1e-9 and ee < 1e-9, (rank, e1, e2, ee)


def _worker_sharded(rank, world, port, out):
    """The sharded dense update (tsvgp.cu::sharded_reduce_and_form_G): statistics reduce-scattered by rows, Y = B K9^-1 and
    G2 = K9^-1 Y on each rank's rows, two all-gathers — against the replicated update and the reference-order step."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import scipy.linalg as sla
    from oracle import tsvgp_oracle as orc
    from tests import algo_model as am
    import tsvgp_b200.synth as synth
    from tsvgp_b200 import shard_rows

    cfg = synth.describe("cfg3")
    N, M, num_data, lr, jitter = 900, 64, 9000, 0.5, 1e-9
    X, Y, Z = synth.make_minibatch(cfg, n_rows=N, M=M)
    kernel, lik = synth.build_objects(cfg, orc)
    ref = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(Z.copy()), num_data=num_data)
    ref.natgrad_step((X, Y), lr=lr)
    l1, L2 = ref.lambda_1.copy(), ref.lambda_2_sqrt[0].copy()
    lo, hi = shard_rows(N, world, rank)
    K = kernel.K(Z)
    pre = am.prepare(K, l1, L2)
    Kuf = kernel.K(Z, X[lo:hi])
    mu, var = am.marginals(Kuf, kernel.K_diag(X[lo:hi]), pre["T"], pre["alpha"])
    _, g, h = lik.ve_and_grads(mu, var[:, None], Y[lo:hi])
    B, b = am.local_statistics(Kuf, g[:, 0], np.minimum(h[:, 0], -1e-8))
    # reduce-scatter by rows (gloo has no reduce_scatter: all_reduce + this rank's slice is the same data movement result)
    R = M // world
    tB = torch.from_numpy(B.copy()); dist.all_reduce(tB)
    Brows = tB.numpy()[rank * R:(rank + 1) * R]
    tb_ = torch.from_numpy(b.copy()); dist.all_reduce(tb_)
    C9 = sla.cholesky(K + jitter * np.eye(M), lower=True)
    C9inv = sla.solve_triangular(C9, np.eye(M), lower=True)
    K9inv = C9inv.T @ C9inv
    Yrows = torch.from_numpy(Brows @ K9inv)
    parts = [torch.zeros_like(Yrows) for _ in range(world)]
    dist.all_gather(parts, Yrows)
    Yfull = torch.cat(parts).numpy()
    Grows = torch.from_numpy(K9inv[rank * R:(rank + 1) * R] @ Yfull)
    parts = [torch.zeros_like(Grows) for _ in range(world)]
    dist.all_gather(parts, Grows)
    G2 = torch.cat(parts).numpy()
    G1 = K9inv @ tb_.numpy()
    scale = num_data / N
    mZ = K @ pre["alpha"][:, 0]
    n1 = (1 - lr) * l1[:, 0] + lr * scale * (G1 - 2.0 * G2 @ mZ)
    P = (1 - lr) * (L2 @ L2.T) + lr * scale * (-2.0 * G2) + jitter * np.eye(M)
    nL2 = -sla.cholesky(np.tril(P) + np.tril(P, -1).T, lower=True)     # the device factors the lower triangle
    ref.natgrad_step((X, Y), lr=lr)
    rel = lambda a, c: float(np.max(np.abs(a - c)) / np.max(np.abs(c)))  # noqa: E731
    out.put((rank, rel(n1, ref.lambda_1[:, 0]), rel(nL2 @ nL2.T, ref.lambda_2[0]), n1.tobytes()))
    dist.destroy_process_group()


def test_two_rank_sharded_update_matches_full_batch():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sharded, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e1, e2, _ in res:
        assert e1 < 1e-9 and e2 < 1e-9, (rank, e1, e2)
    assert res[0][3] == res[1][3]        # every rank ends with the same bits (gathered G2, replicated factorisation)
