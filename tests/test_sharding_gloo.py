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


def _worker(rank, world, port, out):
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

    lo, hi = shard_rows(N, world, rank)
    K = kernel.K(Z)
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
