#!/usr/bin/env python
"""
bench.py — natgrad_step datapoints/s of the B200 t-SVGP path (BASELINE.json metric), one JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg3] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload: BASELINE.json's metric is quoted "at 1/2/4/8 B200", which is configs[2] (cfg3: N=10M, M=2048, D=16, Matern-5/2, minibatch 1M
"sharded over 1/2/4/8 B200"); configs[1] (cfg2) is pinned to one B200 and holds 0.15 ms of FP64 work per step, so it measures launch
latency rather than the path.  cfg3 fits one GPU and is the default; the default N=1 run also measures cfg2 (configs[1]) and reports it
under `other_configs`; `--config cfgK` selects any config as the main line.

A "step" is one `t_SVGP.natgrad_step` (posterior factors + streaming statistics pass + all-reduce + dense site update) over
one minibatch of the named config, sharded by rows over the N ranks (strong scaling: the minibatch is fixed).
  value : minibatch rows / step time with the minibatch resident in HBM when the clock starts (4 distinct minibatches rotate)
  e2e   : the same through the public API with HOST (pinned) buffers: H2D of the rank's rows every step, and the pre-step
          ELBO and lambda_1 read back, inside the timed region
  roofline     : the DMMA weighted-SYRK kernel, timed live with CUDA events (single-stream profiled pass after the timed region)
  cpu_baseline : the NumPy/SciPy restatement of the reference (oracle/) on the host cores, bounded sample, extrapolated in N
`--impl reference` times that CPU restatement alone (TensorFlow/GPflow are not installable in this image).
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FP64_PEAK_TFLOPS = 37.1   # measured on this pool: tools/fp64_peak (DMMA m8n8k4 issue rate), profiles/fp64_peak_r01.txt
N_RESIDENT_MINIBATCHES = 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="cfg3")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--minibatch", type=int, default=None, help="override the config's minibatch rows (debugging)")
    ap.add_argument("--M", type=int, default=None, help="override the number of inducing points (debugging)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-grad", action="store_true", help="skip the (reported, untimed-in-the-metric) M-step gradient measurement")
    ap.add_argument("--keep-cache", action="store_true", help="do not invalidate the kernel-matrix factors every step")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (tsvgp_set_option), repeatable")
    return ap.parse_args()


# ---- clocks -------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
            out, _ = self.p.communicate()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---- CPU baseline (oracle = NumPy restatement of the reference) -------------------------------------------------------
def cpu_reference_step_time(cfg, M, n_rows, repeats=1):
    """Seconds of one reference-order natgrad_step (oracle) on `n_rows` rows with the full M, from non-trivial sites."""
    import tsvgp_b200.synth as synth
    from oracle import tsvgp_oracle as orc
    X, Y, Z = synth.make_minibatch(cfg, n_rows=n_rows, M=M)
    kernel, lik = synth.build_objects(cfg, orc)
    m = orc.OracleTSVGP(kernel, lik, orc.InducingPoints(Z), num_data=cfg["N"])
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        m.natgrad_step((X, Y), lr=cfg["lr"])
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


CPU_FLOP_BUDGET = 2.5e12   # ~10-30 s of host work at a few hundred GFLOP/s
CPU_M_CAP = 4096           # above this the 37 M^3 dense part alone exceeds the budget: measure at the cap and scale


def reference_flops(n, M):
    return 8.0 * n * M * M + 37.0 * M ** 3   # what the reference's operation order executes


def cpu_fit(cfg, M, Nb, budget=None):
    """Reference-order step time on the host for the (Nb, M) workload from a bounded sample -> (seconds, description)."""
    def timer(n, m):
        return cpu_reference_step_time(cfg, m, n)
    budget = budget or CPU_FLOP_BUDGET
    if reference_flops(Nb, M) <= budget:
        timer(min(Nb, 64), M)
        t = min(timer(Nb, M), timer(Nb, M))
        return t, f"the full {Nb}-row minibatch at M={M} (best of 2)"
    Mc = min(M, CPU_M_CAP)
    n_large = int(max(1024, (budget - 37.0 * Mc ** 3) / (8.0 * Mc * Mc) / 1.25)) // 256 * 256
    n_large = int(min(max(n_large, 2048), 65536, Nb))
    n_small = max(256, n_large // 4)
    timer(64, Mc)                                  # warm the BLAS threads and the allocator
    t_small, t_large = timer(n_small, Mc), timer(n_large, Mc)
    b = max((t_large - t_small) / (n_large - n_small), 1e-9)
    a = max(t_small - b * n_small, 0.0)
    note = f"t(N)=a+bN fitted at {n_small} and {n_large} rows, M={Mc} (a={a:.3f}s dense part, b={b * 1e6:.2f}us/row)"
    if Mc != M:
        a *= (M / Mc) ** 3
        b *= (M / Mc) ** 2
        note += (f", then scaled to M={M} by (M/{Mc})^3 for a and (M/{Mc})^2 for b because the 37 M^3 dense flops at M={M} alone "
                 "exceed the sample budget")
    return a + b * Nb, note + f", extrapolated to the {Nb}-row minibatch"


def cpu_baseline(cfg, M, Nb):
    t_full, note = cpu_fit(cfg, M, Nb)
    return {"value": Nb / t_full, "unit": "datapoints/s", "cores": cpu_threads(), "kind": "port",
            "sample": "oracle (NumPy/SciPy restatement of the reference, not TensorFlow) natgrad_step: " + note,
            "host_cpus": os.cpu_count(), "t_full_step_s": t_full}


# ---- the reference arm -----------------------------------------------------------------------------------------------
def run_reference(args, cfg, M, Nb):
    """The reference's CPU implementation of the path (oracle port: the reference itself needs GPflow/TensorFlow, which this
    image cannot install).  Each step is one bounded sample (see cpu_fit); the value is the mean over the timed steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, note = [], ""
    per_step_budget = max(CPU_FLOP_BUDGET * 6.0 / max(args.warmup + args.steps, 1), 1.2 * 37.0 * min(M, CPU_M_CAP) ** 3)
    for i in range(args.warmup + args.steps):   # the whole run stays within a few minutes of host time
        t_full, note = cpu_fit(cfg, M, Nb, budget=min(per_step_budget, CPU_FLOP_BUDGET))
        if i >= args.warmup:
            vals.append(t_full)
    t_full = float(np.mean(vals))
    value = Nb / t_full
    line = {
        "impl": "reference", "metric": "natgrad_step datapoints/sec", "value": value, "unit": "datapoints/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_full * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(cfg, M, Nb, args.gpus),
        "cpu_baseline": {"value": value, "unit": "datapoints/s", "cores": cpu_threads(), "kind": "port",
                         "sample": "each step = oracle (NumPy/SciPy restatement; GPflow/TensorFlow not installable here) natgrad_step: " + note,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": "datapoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(cfg, M, Nb, n_gpus):
    return {"workload": f"{cfg['name']}: {cfg['lik']} likelihood, synthetic D={cfg['D']}, N={cfg['N']}, M={M}, {cfg['kernel']} kernel, "
                        f"minibatch {Nb} sharded by rows over {n_gpus} GPU(s), lr={cfg['lr']}",
            "M": M, "D": cfg["D"], "minibatch": Nb, "num_data": cfg["N"], "kernel": cfg["kernel"], "likelihood": cfg["lik"],
            "parallelism": f"rows/{n_gpus} + 1 allreduce, dense phase replicated",
            "l2": f"inputs larger than L2: {N_RESIDENT_MINIBATCHES} distinct resident minibatches rotate; the Kuf slabs are L2-resident by design"}


# ---- our arm -------------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    import tsvgp_b200.synth as synth
    cfg = synth.describe(args.config)
    M = args.M or cfg["M"]
    Nb = args.minibatch or cfg["Nb"]
    if args.impl == "reference":
        return run_reference(args, cfg, M, Nb)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    # stdout carries exactly ONE line (the JSON): keep NCCL's version banner (printed to stdout when the box exports
    # NCCL_DEBUG=VERSION/INFO) out of it
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE") and not os.environ.get("BENCH_KEEP_NCCL_DEBUG"):
        os.environ["NCCL_DEBUG"] = "WARN"
    if world > 1:
        import torch.distributed as dist   # plumbing only: rendezvous, barrier, max over ranks (gloo; no tensors on the GPU)
        dist.init_process_group("gloo")

    import tsvgp_b200 as tb
    from tsvgp_b200 import standins as st

    kernel, lik = synth.build_objects(cfg, st)
    lo, hi = tb.shard_rows(Nb, world, rank)
    n_local = hi - lo
    # synthetic minibatches (same seeds on every rank; each rank keeps its rows)
    mbs_host = []
    Z = None
    for i in range(N_RESIDENT_MINIBATCHES):
        X, Y, Zi = synth.make_minibatch(cfg, n_rows=Nb, M=M, seed_offset=100 * i)
        if Z is None:
            Z = Zi
        mbs_host.append((np.ascontiguousarray(X[lo:hi]), np.ascontiguousarray(Y[lo:hi])))
        del X, Y

    model = tb.t_SVGP(kernel, lik, Z, num_data=cfg["N"], device=local_rank)
    for kv in args.opt:
        k, v = kv.split("=")
        model.set_option(k, float(v))
    if world > 1:
        import torch
        if rank == 0:
            uid = tb.comm_unique_id()
            t = torch.tensor(list(uid), dtype=torch.uint8)
        else:
            t = torch.zeros(128, dtype=torch.uint8)
        dist.broadcast(t, 0)
        sys.stdout.flush()
        saved = os.dup(1)              # NCCL may printf its version banner to stdout during ncclCommInitRank: send it to stderr
        os.dup2(2, 1)
        try:
            model.init_comm(world, rank, bytes(t.tolist()))
        finally:
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    mbs_dev = [(model.device_array(X), model.device_array(Y)) for X, Y in mbs_host]
    invalidate = not args.keep_cache

    def step_resident(i):
        if invalidate:
            model.set_option("invalidate", 1)   # kernel matrices and their factors are rebuilt every step, as the reference does
        model.set_data(mbs_dev[i % N_RESIDENT_MINIBATCHES])
        model.natgrad_step(lr=cfg["lr"], global_minibatch_size=Nb)

    # ---------- value: device-resident inputs ----------
    for i in range(args.warmup):
        step_resident(i)
    model.sync(); barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches = 0
    phase = {"prepare": 0.0, "stream": 0.0, "allreduce": 0.0, "dense": 0.0}
    model.timer_start()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_resident(args.warmup + i)
        tm = model.timings()
        launches += int(tm["launches"])
        for k in phase:
            phase[k] += tm[k] / args.steps
    ms_dev = model.timer_stop()
    wall = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = max_over_ranks(ms_dev)
    ms_per_step = ms_total / args.steps
    value = Nb / (ms_per_step * 1e-3)
    route = model.timings()

    # ---------- e2e: host (pinned) buffers through the public API, H2D + result read-back every step ----------
    e2e = None
    if not args.no_e2e:
        pinned = []
        for X, Y in mbs_host:
            px, py = tb.pinned_empty(X.shape), tb.pinned_empty(Y.shape)
            px[...] = X; py[...] = Y
            pinned.append((px, py))

        # the public API with host buffers: the input pipeline (tsvgp_b200.stream_minibatches pattern) copies minibatch i + 1 from
        # pinned host memory while step i computes; every step's rows cross the PCIe bus inside the timed region, and the
        # pre-step ELBO and lambda_1 are read back every step
        def run_e2e(n_steps):
            model.stage_data(pinned[0])
            for i in range(n_steps):
                model.commit_staged()
                model.stage_data(pinned[(i + 1) % N_RESIDENT_MINIBATCHES])
                if invalidate:
                    model.set_option("invalidate", 1)
                e = model.natgrad_step(lr=cfg["lr"], global_minibatch_size=Nb, return_elbo=True)
                l1 = model.lambda_1
            model.commit_staged()   # drain the last prefetch
            return e, l1

        run_e2e(min(args.warmup, 2))
        model.sync(); barrier()
        model.timer_start()
        run_e2e(args.steps)
        ms_e2e = max_over_ranks(model.timer_stop()) / args.steps
        barrier()
        e2e = {"value": Nb / (ms_e2e * 1e-3), "unit": "datapoints/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(n_local * (cfg["D"] + 1) * 8), "d2h_bytes_per_step": int(8 + 8 * M),
               "api": "stage_data(next pinned (X, Y)) / commit_staged + t_SVGP.natgrad_step(return_elbo=True) + .lambda_1, per rank on its rows"}

    # ---------- M-step gradient pass (next-row feature; reported, not part of the metric) ----------
    grad_ms = None
    if world == 1 and cfg["D"] <= 63 and not args.no_grad:
        model.set_data(mbs_dev[0])
        model.elbo_and_grad(global_minibatch_size=Nb)
        model.timer_start()
        for i in range(2):
            model.elbo_and_grad(global_minibatch_size=Nb)
        grad_ms = model.timer_stop() / 2

    # ---------- roofline: per-kernel CUDA-event timing, single stream ----------
    model.set_option("profile", 1)
    step_resident(0); step_resident(1)
    kp = model.kernel_profile()
    model.set_option("profile", 0)
    Mp = (M + 127) // 128 * 128
    syrk_ms, syrk_n = kp["syrk"]
    var_ms, var_n = kp["variance_gemm"]
    rows_per_launch = n_local / max(syrk_n, 1)
    syrk_flops = float(M) * M * rows_per_launch   # algorithmic: lower triangle of k k^T, 2 flops per entry = M^2 per point
    syrk_t = syrk_ms * 1e-3 / max(syrk_n, 1)
    achieved = syrk_flops / syrk_t / 1e12
    traffic = None
    summ = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(summ):
        try:
            traffic = json.load(open(summ)).get(args.config, {}).get("syrk_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "tensor", "kernel": "gemm_kernel<kc,kc,scale> (weighted SYRK B += K diag(h) K^T, DMMA m8n8k4)",
                "achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS,
                "traffic": traffic,
                "peak_source": "measured FP64 DMMA issue rate on this pool (tools/fp64_peak -> profiles/fp64_peak_r01.txt); "
                               "MEASURED_PEAKS.json carries no FP64 figure and the profiling guide states no FP64 fallback",
                "flops_per_launch": syrk_flops, "avg_launch_ms": syrk_t * 1e3, "launches_profiled": syrk_n}
    var_t = var_ms * 1e-3 / max(var_n, 1)
    q = lik_q(cfg)
    f_pt = synth.flops_per_point(M, cfg["D"], q)
    step_flops = Nb * f_pt + synth.dense_flops(M)
    kernels = {k: {"ms_total": v[0], "launches": v[1]} for k, v in kp.items()}
    kernels["variance_gemm"]["achieved_tflops"] = float(M) * M * rows_per_launch / var_t / 1e12 if var_t > 0 else None
    extra = {
        "step_fp64_frac": step_flops / (ms_per_step * 1e-3) / (world * FP64_PEAK_TFLOPS * 1e12),
        "step_algorithmic_flops": step_flops, "flops_per_point": f_pt,
        "hbm_stream_frac": Nb * 8 * (cfg["D"] + 1) / max(phase["stream"] * 1e-3, 1e-9) / (world * 6554.2e9),
        "phase_ms": phase, "kernels": kernels, "route": int(route["route"]), "cond_est": route["cond_est"],
        "wall_ms_per_step": wall * 1e3 / args.steps,
        "elbo_and_grad_ms": grad_ms,
    }

    line = {
        "metric": "natgrad_step datapoints/sec", "value": value, "unit": "datapoints/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(cfg, M, Nb, world), "clocks": clocks, "e2e": e2e,
        "gpu_launches": launches, "roofline": roofline, "detail": extra,
    }
    if rank == 0 and world == 1 and args.config == "cfg3" and args.minibatch is None and args.M is None and not args.no_e2e:
        line["other_configs"] = {"cfg2": quick_config(tb, st, synth, "cfg2", local_rank)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(cfg, M, Nb)
    if rank == 0:
        print(json.dumps(line), flush=True)
    model.close()
    if dist is not None:
        dist.destroy_process_group()


def quick_config(tb, st, synth, name, device):
    """configs[1] measured in the same run: natgrad_step datapoints/s, device-resident and end to end (pinned host buffers)."""
    cfg = synth.describe(name)
    M, Nb = cfg["M"], cfg["Nb"]
    kernel, lik = synth.build_objects(cfg, st)
    X, Y, Z = synth.make_minibatch(cfg)
    m = tb.t_SVGP(kernel, lik, Z, num_data=cfg["N"], device=device)
    xd, yd = m.device_array(X), m.device_array(Y)
    px, py = tb.pinned_empty(X.shape), tb.pinned_empty(Y.shape)
    px[...] = X; py[...] = Y

    def resident():
        m.set_option("invalidate", 1)
        m.set_data((xd, yd))
        m.natgrad_step(lr=cfg["lr"])

    def e2e():
        m.set_option("invalidate", 1)
        e = m.natgrad_step((px, py), lr=cfg["lr"], return_elbo=True)
        return e, m.lambda_1

    out = {}
    for key, fn in (("value", resident), ("e2e", e2e)):
        for _ in range(5):
            fn()
        m.sync()
        m.timer_start()
        for _ in range(30):
            fn()
        ms = m.timer_stop() / 30
        out[key] = Nb / (ms * 1e-3)
        out[key + "_ms_per_step"] = ms
    q = lik_q(cfg)
    out.update(unit="datapoints/s", workload=f"{name}: {cfg['lik']} GH-20, D={cfg['D']}, M={M}, minibatch {Nb} (BASELINE configs[1])",
               step_fp64_frac=(Nb * synth.flops_per_point(M, cfg["D"], q) + synth.dense_flops(M)) / (out["value_ms_per_step"] * 1e-3) / (FP64_PEAK_TFLOPS * 1e12))
    m.close()
    return out


def lik_q(cfg):
    return 0 if cfg["lik"] == "Gaussian" else 20


if __name__ == "__main__":
    main()
