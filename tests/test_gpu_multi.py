"""
GPU, >= 2 devices: the minibatch sharded by rows over two ranks (one context per GPU, one NCCL all-reduce of the statistics)
gives the single-GPU result.  Skipped on a one-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`.
"""
import multiprocessing as mp
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30)
        return len([l for l in out.stdout.splitlines() if l.startswith("GPU ")])
    except Exception:
        return 0


def _case(M=300):
    import tsvgp_b200.synth as synth
    from tsvgp_b200 import standins as st
    cfg = synth.describe("cfg5")
    X, Y, Z = synth.make_minibatch(cfg, n_rows=5001, M=M)
    kernel, lik = synth.build_objects(cfg, st)
    return cfg, X, Y, Z, kernel, lik


def _rank(rank, world, conn, out, opts, M):
    sys.path.insert(0, ROOT)
    import tsvgp_b200 as tb
    cfg, X, Y, Z, kernel, lik = _case(M)
    m = tb.t_SVGP(kernel, lik, Z, num_data=50_010, device=rank)
    steps = opts.get("_steps", 2)
    for k, v in opts.items():
        if not k.startswith("_"):
            m.set_option(k, v)
    if rank == 0:
        uid = tb.comm_unique_id()
        for c in conn:
            c.send(uid)
    else:
        uid = conn.recv()
    m.init_comm(world, rank, uid)
    lo, hi = tb.shard_rows(X.shape[0], world, rank)
    elbos = []
    for _ in range(steps):
        if opts.get("_invalidate"):       # as after an M-step: the kernel matrix and both factorisation chains are rebuilt
            m.set_option("invalidate", 1)
        elbos.append(m.natgrad_step((X[lo:hi], Y[lo:hi]), lr=cfg["lr"], global_minibatch_size=X.shape[0], return_elbo=True))
    role = m.timings()["chain_role"]   # of the last step: 0 both chains here, 1 / 2 posterior factors / K9 chain built here, the other received
    # predict_f is not a collective.  Right after a step the posterior factors are stale on every rank; rank 0 ALONE
    # predicts (it must rebuild them without an all-reduce even when the dense products are distributed), then every rank calls
    # the collective elbo() — rank 0's privately rebuilt cache must not unbalance the other ranks' collectives.
    mu_solo = m.predict_f(X[50:100])[0] if rank == 0 else None
    elbos.append(m.elbo((X[lo:hi], Y[lo:hi]), global_minibatch_size=X.shape[0]))
    mu, var = m.predict_f(X[:50])
    out.put((rank, m.lambda_1, m.lambda_2, elbos, mu, var, mu_solo, role))
    m.close()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("opts,M", [({"shard_min_m": 1 << 30}, 300), ({"dist_min_m": 128, "shard_min_m": 1 << 30}, 300), ({"shard_min_m": 128}, 500),
                                    ({"dist_min_m": 128, "shard_min_m": 128}, 500),
                                    ({"shard_min_m": 128, "split_chains": 1, "_invalidate": 1, "_steps": 3}, 500),
                                    ({"shard_min_m": 128, "split_chains": 0, "_invalidate": 1, "_steps": 3}, 500)],
                         ids=["replicated_dense", "distributed_dense", "sharded_update", "sharded_update_with_distributed_products",
                              "chains_split_over_the_pair", "chains_side_by_side"])
def test_two_gpu_sharded_step_matches_single_gpu(opts, M):
    # distributed_dense: the M x M products of the dense phase are dealt out row-cyclically over the ranks and assembled by
    # all-reduce (forced here at small M; by default from M >= 4096)
    # sharded_update: statistics reduce-scattered by tile rows, G2 = K9^-1 B K9^-1 formed on each rank's rows, two all-gathers
    # (forced here at M = 500 = 4 tile rows; by default from M >= 2048 when the tile rows divide by the ranks)
    # chains_split_over_the_pair: from the second step on (kernel matrix invalidated before every step, route speculated) rank 0
    # builds the posterior factors, rank 1 the Kuu + jitter I chain, and each broadcasts its result — same bits as side by side
    import tsvgp_b200 as tb
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    a, b = ctx.Pipe()
    procs = [ctx.Process(target=_rank, args=(0, 2, [a], out, opts, M)), ctx.Process(target=_rank, args=(1, 2, b, out, opts, M))]
    for p in procs:
        p.start()
    res = sorted([out.get(timeout=300) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg, X, Y, Z, kernel, lik = _case(M)
    single = tb.t_SVGP(kernel, lik, Z, num_data=50_010)
    steps = opts.get("_steps", 2)
    e = [single.natgrad_step((X, Y), lr=cfg["lr"], return_elbo=True) for _ in range(steps)] + [None]
    e[steps] = single.elbo((X, Y))
    mu, var = single.predict_f(X[:50])
    rel = lambda x, y: float(np.max(np.abs(np.asarray(x) - np.asarray(y))) / np.max(np.abs(y)))  # noqa: E731
    # summation order only (1e-11); the sharded update forms G2 = K9^-1 (B K9^-1) instead of (K9^-1 B) K9^-1: the two orders differ
    # by rounding amplified by cond(Kuu), so it is held to a tenth of the parity tolerance
    tol = 1e-10 if "shard_min_m" in opts and opts["shard_min_m"] < 1 << 20 else 1e-11
    for rank, l1, l2, elbos, mu_r, var_r, mu_solo, role in res:
        assert role == ((1 + rank) if opts.get("_invalidate") and opts.get("split_chains", 0) else 0)
        assert rel(l1, single.lambda_1) < tol and rel(l2, single.lambda_2) < tol
        assert rel(elbos, e) < tol and rel(mu_r, mu) < tol and rel(var_r, var) < tol
        if rank == 0:
            assert rel(mu_solo, single.predict_f(X[50:100])[0]) < tol
    np.testing.assert_array_equal(res[0][1], res[1][1])   # replicated dense phase: both ranks hold identical sites
    np.testing.assert_array_equal(res[0][2], res[1][2])
    single.close()
