#!/usr/bin/env python
"""bench.py — SAFER2 epoch on the synthetic ML-20M shape at d=256 (BASELINE.json configs[2]).

    python bench.py --gpus N --steps K --warmup W           # our CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W   # the reference's CPU path (oracle port), rank 0 only

A "step" is one Train() epoch (safer2.h:266-334): z update, user half-step, weighted
Gramian, item half-step, item Gramian, per-user loss, xi Newton iterations.  `value` is
row-solves/s = (users + items with history) / epoch seconds, inputs resident in HBM;
`e2e` is the same metric through the C ABI with the factors in pinned HOST memory
(H2D of U,V before and D2H of U,V + scalars after every epoch).  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SHAPES = {
    # name: (num_users, num_items, nnz)  — SURVEY.md 8d
    "ml20m": (138493, 26744, 20000263),
    "ml1m": (6040, 3706, 1000209),
    "msd": (571355, 41140, 33633450),
    "tiny": (3000, 1500, 150000),
}
# README.md:79 (ML-20M SAFER2 command line)
SAFER2_ML20M = dict(model="safer2", uobs_weight=0.002, alpha=0.3, reg=0.002, stdev=0.1, bandwidth=0.18,
                    pd_iterations=1, xi_iterations=5, use_snr=1, sampling_ratio=0.1, snr_seed=1)


def synth_interactions(num_users, num_items, nnz, seed=98765):
    """Seeded synthetic interactions with the ML-20M-like statistics of SURVEY.md 8d:
    log-normal history sizes (sd 0.97, min 5), Zipf-like item popularity, no duplicate
    pairs, grouped by user in random user order, items unsorted within a user."""
    rng = np.random.default_rng(seed)
    zn = rng.normal(0.0, 0.97, num_users)
    lo, hi = 0.0, 12.0
    for _ in range(60):
        mu = 0.5 * (lo + hi)
        n = np.clip(np.round(np.exp(mu + zn)), 5, num_items // 4)
        if n.sum() > nnz:
            hi = mu
        else:
            lo = mu
    n_u = np.clip(np.round(np.exp(lo + zn)), 5, num_items // 4).astype(np.int64)
    pop = 1.0 / (np.arange(num_items) + 20.0) ** 0.9
    cdf = np.cumsum(pop / pop.sum())
    item_of_rank = rng.permutation(num_items).astype(np.int64)
    m_u = (n_u * 1.35).astype(np.int64) + 16
    users = np.repeat(np.arange(num_users, dtype=np.int64), m_u)
    draws = np.searchsorted(cdf, rng.random(users.shape[0]))
    np.minimum(draws, num_items - 1, out=draws)
    key = np.unique(users * num_items + item_of_rank[draws])
    users = key // num_items
    items = key % num_items
    del key, draws
    prio = rng.random(users.shape[0])
    user_rank = rng.permutation(num_users)
    order = np.lexsort((prio, user_rank[users]))
    users, items = users[order], items[order]
    del order, prio
    start = np.flatnonzero(np.r_[True, users[1:] != users[:-1]])
    seg = np.repeat(start, np.diff(np.r_[start, users.shape[0]]))
    pos = np.arange(users.shape[0]) - seg
    keep = pos < n_u[users]
    return users[keep].astype(np.int32), items[keep].astype(np.int32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                               parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline_sample(args, users, items, num_users, num_items, cfg, frac=0.01, repeats=1):
    """The oracle port of the reference's CPU path on a bounded sample of the SAME workload:
    the user half-step + loss on the full histories of `frac` of the users, and the item
    half-step on the full histories of `frac` of the items (same mix of row solves as a full
    epoch).  Returns row-solves/s with all host threads (the reference's thread model)."""
    from oracle import loader as O
    rng = np.random.default_rng(4242)
    su = np.sort(rng.choice(num_users, max(1, int(num_users * frac)), replace=False))
    si = np.sort(rng.choice(num_items, max(1, int(num_items * frac)), replace=False))
    mu = np.isin(users, su)
    mi = np.isin(items, si)
    dsA = O.Dataset.from_tuples(users[mu], items[mu])
    dsB = O.Dataset.from_tuples(users[mi], items[mi])
    m = O.Model(num_users, num_items, init_seed=12345, **cfg)
    m.initialize(dsA)
    m.initialize(dsB)  # finite history sizes for every user the sampled items touch (timing only)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        m.stage(dsA, 1)   # StepU on sampled users
        m.stage(dsA, 4)   # ComputeUserLoss on sampled users
        t1 = time.perf_counter()
        m.stage(dsB, 2)   # StepV on sampled items
        m.stage(dsB, 3)   # item Gramian (full V)
        t2 = time.perf_counter()
        dt = t2 - t0
        best = dt if best is None else min(best, dt)
    rows = dsA.distinct_users + dsB.distinct_items
    return {"value": rows / best, "unit": "row-solves/s", "cores": O.num_threads(), "kind": "port",
            "seconds": best, "rows": rows,
            "sample": f"{frac:.0%} of users (StepU+ComputeUserLoss, {dsA.num_tuples} tuples) + {frac:.0%} of items "
                      f"(StepV, {dsB.num_tuples} tuples) of the same workload; Eigen-free CPU restatement of the "
                      f"reference (Eigen unavailable in image), -O3 -march=native, {O.num_threads()} threads"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="ml20m", choices=sorted(SHAPES))
    ap.add_argument("--dim", type=int, default=256)
    ap.add_argument("--model", default="safer2")
    ap.add_argument("--cpu-frac", type=float, default=0.15)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-stages", action="store_true", help="print per-stage times to stderr")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    num_users, num_items, nnz = SHAPES[args.shape]
    cfg = dict(SAFER2_ML20M)
    cfg.update(model=args.model, dim=args.dim)
    config = {"workload": f"SAFER2 d={args.dim} synthetic {args.shape} shape ({num_users} users x {num_items} items, "
                          f"~{nnz} nnz), use_snr=1 sampling_ratio=0.1 (BASELINE.json configs[2], README.md:79 flags)",
              "model_name": args.model, "dim": args.dim, "shape": args.shape,
              "l2_cache": "inputs_exceed_l2 (CSR+factors ~650 MB vs 126 MB L2); no explicit flush",
              "parallelism": f"row-sharded dp{world}" if world > 1 else "single GPU"}

    if args.impl == "reference":
        if rank != 0:
            return
        t_gen = time.time()
        users, items = synth_interactions(num_users, num_items, nnz)
        # one "step" = the bounded sample of cpu_baseline_sample; warmup+steps repeats of it
        res = None
        vals = []
        for i in range(args.warmup + args.steps):
            res = cpu_baseline_sample(args, users, items, num_users, num_items, cfg, frac=args.cpu_frac)
            if i >= args.warmup:
                vals.append(res["value"])
        v = float(np.mean(vals))
        res["value"] = v
        out = {"impl": "reference", "metric": "safer2_epoch_row_solves_per_s", "value": v, "unit": "row-solves/s",
               "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": 1e3 * res["rows"] / v, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
               "cpu_baseline": res,
               "e2e": {"value": v, "unit": "row-solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gen_seconds": time.time() - t_gen}
        print(json.dumps(out))
        return

    import torch
    import helpers
    pkg = helpers.load_pkg()
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = pkg.Context(local_rank, stream=stream.cuda_stream)
    if world > 1:
        uid = [pkg.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.init_comm(rank, world, uid[0])

    users, items = synth_interactions(num_users, num_items, nnz)
    ds = pkg.Dataset(ctx, users, items)
    m = pkg.Model(ctx, num_users, num_items, **cfg)
    m.init_factors(12345)
    m.initialize(ds)
    ctx.sync()
    rows_per_epoch = cfg["pd_iterations"] * (ds.distinct_users + ds.distinct_items)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # per-stage CUDA events (on the launching stream) are recorded INSIDE the timed region; every epoch
    # re-records them, so what is read back after the loop is the last timed epoch
    ctx.set_profiling(1)
    for _ in range(args.warmup):
        m.train(ds)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        m.train(ds)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    stage_ms = {}
    for name, t in ctx.stage_times():
        stage_ms[name] = stage_ms.get(name, 0.0) + float(t)   # a name that occurs twice (all-gathers) is summed
    ctx.set_profiling(0)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = rows_per_epoch / (ms_per_step * 1e-3)
    sc = m.scalars()

    # ---- end-to-end through the C ABI with host buffers --------------------------------
    d = args.dim
    Uh = torch.empty((num_users, d), dtype=torch.float32, pin_memory=True).numpy()
    Vh = torch.empty((num_items, d), dtype=torch.float32, pin_memory=True).numpy()
    m.factors(Uh, Vh)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for _ in range(args.steps):
        # H2D of this epoch's inputs from pinned host memory and D2H of the updated factors; with several
        # ranks each one moves the rows it owns (the blocks are all-gathered over NVLink after the upload)
        m.upload_factors_sharded(ds, Uh, Vh)
        m.train_to_host(ds, Uh, Vh)    # synchronises; the D2H of U runs under the item half-step
        sc = m.scalars()              # D2H of xi / weighted loss / mean z
    f1.record(stream)
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = rows_per_epoch / (e2e_ms / args.steps * 1e-3)
    fbytes = 4 * d * (num_users + num_items)

    # ---- roofline from the per-stage device times of the last timed epoch ------
    n_tuples = ds.num_tuples
    # half-step bytes = nnz*(4d+4) + 4*(R+1) + 4*R+*d (+4*nnz weights on the item side)
    bytes_u = n_tuples * (4 * d + 4) + 4 * (num_users + 1) + 4 * ds.distinct_users * d
    bytes_v = n_tuples * (4 * d + 4) + 4 * (num_items + 1) + 4 * ds.distinct_items * d + 4 * n_tuples
    t_rows = (stage_ms.get("step_U", 0.0) + stage_ms.get("step_V", 0.0)) * 1e-3
    peak, peak_src = measured_peaks()
    achieved = (bytes_u + bytes_v) / world / t_rows / 1e9 if t_rows > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "row_solve (step_U + step_V launches: CSR gather + SYRK + Cholesky)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src,
                # dram__bytes_read+write of the epoch's row-kernel launches from one `ncu --set full` capture of this
                # workload on one GPU (profiles/r01_row_solve_tc_ncu_summary.txt); not re-measured per run
                "traffic": 5.79e9 if (args.shape == "ml20m" and args.dim == 256 and world == 1
                                      and cfg["model"] == "safer2") else None,
                "algorithmic_bytes_per_epoch": bytes_u + bytes_v,
                "kernel_share_of_step": t_rows * 1e3 / max(1e-9, sum(stage_ms.values())),
                "stage_ms": stage_ms}
    if args.profile_stages and rank == 0:
        print(json.dumps(stage_ms, indent=1), file=sys.stderr)

    if rank != 0:
        return
    out = {"metric": "safer2_epoch_row_solves_per_s", "value": value, "unit": "row-solves/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "epoch_s": ms_per_step * 1e-3,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": config, "clocks": clocks, "gpu_launches": int(launches),
           "e2e": {"value": e2e_value, "unit": "row-solves/s", "h2d_bytes_per_step": fbytes,
                   "d2h_bytes_per_step": fbytes + 12 * world, "ms_per_step": e2e_ms / args.steps,
                   "note": "whole-job bytes: every rank moves its own row shard of U and V"},
           "roofline": roofline,
           "check": {"xi": sc["xi"], "weighted_loss": sc["weighted_loss"], "mean_weight": sc["mean_weight"],
                     "rows_per_epoch": rows_per_epoch, "num_tuples": n_tuples}}
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_sample(args, users, items, num_users, num_items, cfg, frac=args.cpu_frac)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
