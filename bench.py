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
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1380.0),
                "source": "measured (MEASURED_PEAKS.json)", "measured": True}
    return {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1380.0, "source": "fallback (B200_PROFILING.md)",
            "measured": False}


def profile_traffic(args, world, cfg):
    """dram__bytes_read + dram__bytes_write of the epoch's row-kernel launches from the committed `ncu --set full`
    capture of this workload (profiles/r02_row_kernels_dram.json, written by tools/capture_profiles.sh); None when
    this run is a different workload or the capture is absent -- never a hard-coded figure."""
    path = os.path.join(ROOT, "profiles", "r02_row_kernels_dram.json")
    if not (args.shape == "ml20m" and args.dim == 256 and world == 1 and cfg["model"] == "safer2"):
        return None, None
    try:
        d = json.load(open(path))
        return float(d["dram_bytes_per_epoch"]), "profiles/r02_row_kernels_dram.json"
    except (OSError, KeyError, ValueError):
        return None, None


def epoch_flops(n_tuples, rows_u, rows_i, num_users, num_items, d):
    """Flops of one SAFER2 epoch by the REFERENCE's algorithm: lower-triangle SYRK + rhs of both
    half-steps, one d x d Cholesky solve per row, the loss pass and the two Gramians."""
    syrk = 2 * n_tuples * (d * (d + 1) + 2 * d)
    chol = (rows_u + rows_i) * (d ** 3 / 3 + 2 * d * d)
    loss = 2 * n_tuples * d + 3 * n_tuples + num_users * (2 * d * d + 2 * d)
    gram = 2 * (2 * num_users + num_items) * d * d   # U^T diag(z) U, V^T V (cached + recomputed once)
    return {"syrk": syrk, "cholesky": chol, "loss": loss, "gramian": gram, "total": syrk + chol + loss + gram}


def cpu_baseline_sample(args, users, items, num_users, num_items, cfg, frac=0.01, repeats=1):
    """The oracle port of the reference's CPU path on a bounded sample of the SAME workload, with all host
    threads (the reference's thread model): the row-wise stages (StepU + ComputeUserLoss, StepV) run on the full
    histories of `frac` of the users / items and are scaled by the row counts; the stages that do not shard by
    row (the weighted Gramian U^T diag(z) U over all users, V^T V, the xi Newton iterations, the z update) run at
    full size.  Returns the epoch-equivalent row-solves/s.  Built with -DORACLE_FAST (register-blocked SYRK,
    right-looking blocked Cholesky: the shape of Eigen's kernels; timing-only, see oracle/frecsys_oracle.hpp)."""
    os.environ["FRECSYS_ORACLE_FAST"] = "1"   # before the first import of the loader in this process
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
    rows_u_all = int(np.unique(users).size)
    rows_i_all = int(np.unique(items).size)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        m.stage(dsA, 0)   # z update (all users)
        t_z = time.perf_counter() - t0
        t0 = time.perf_counter()
        m.stage(dsA, 1)   # StepU on the sampled users
        m.stage(dsA, 4)   # ComputeUserLoss on the sampled users
        t_u = time.perf_counter() - t0
        U, _ = m.factors()
        t0 = time.perf_counter()
        Gz = O.gramian(U, m.state()["z"])   # U^T diag(z) U over ALL users (safer2.h:504-509), full size
        t_gz = time.perf_counter() - t0
        m.set_gz_override(Gz)
        t0 = time.perf_counter()
        m.stage(dsB, 2)   # StepV on the sampled items (Gramian supplied above)
        t_v = time.perf_counter() - t0
        t0 = time.perf_counter()
        m.stage(dsB, 3)   # V^T V, full size
        t_gv = time.perf_counter() - t0
        t0 = time.perf_counter()
        m.stage(dsA, 5)   # xi Newton / Armijo iterations over all (sub-sampled) user losses
        t_xi = time.perf_counter() - t0
        t_sample = t_z + t_u + t_gz + t_v + t_gv + t_xi
        epoch_equiv = (t_u * rows_u_all / max(1, dsA.distinct_users) + t_v * rows_i_all / max(1, dsB.distinct_items)
                       + t_z + t_gz + t_gv + t_xi)
        if best is None or epoch_equiv < best[0]:
            best = (epoch_equiv, t_sample, dict(z=t_z, step_U_loss=t_u, gramian_Uz=t_gz, step_V=t_v, gramian_V=t_gv, xi=t_xi))
    epoch_equiv, t_sample, parts = best
    rows = rows_u_all + rows_i_all
    fl = epoch_flops(int(users.shape[0]), rows_u_all, rows_i_all, num_users, num_items, cfg["dim"])
    cores = O.num_threads()
    return {"value": rows / epoch_equiv, "unit": "row-solves/s", "cores": cores, "kind": "port",
            "epoch_equivalent_s": epoch_equiv, "sample_seconds": t_sample, "sample_stage_seconds": parts,
            "gflops_per_core": fl["total"] / epoch_equiv / 1e9 / cores,
            "sample": f"{frac:.0%} of users (StepU+ComputeUserLoss, {dsA.num_tuples} tuples) + {frac:.0%} of items "
                      f"(StepV, {dsB.num_tuples} tuples) scaled by row count, plus the full-size weighted Gramian, "
                      f"V^T V, z update and xi iterations; Eigen-free CPU restatement of the reference (Eigen is "
                      f"not in the image) in its timing build (-O3 -march=native -DORACLE_FAST: register-blocked "
                      f"SYRK, right-looking blocked Cholesky), {cores} threads. A stated baseline, not Eigen itself."}


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
    config = {"workload": f"{args.model.upper()} d={args.dim} synthetic {args.shape} shape ({num_users} users x {num_items} items, "
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
        rows_all = int(np.unique(users).size + np.unique(items).size)
        out = {"impl": "reference", "metric": f"{args.model}_epoch_row_solves_per_s", "value": v, "unit": "row-solves/s",
               "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
               # epoch-equivalent time (the bounded sample scaled to the full workload), not the sample's own time
               "ms_per_step": 1e3 * rows_all / v, "sample_ms_per_step": 1e3 * res["sample_seconds"],
               "higher_is_better": True, "scaling": "strong",
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
    # Algorithmic bytes / flops are the per-unit figures of SURVEY.md 8d (restated in DESIGN.md section 5) times
    # the units one epoch processes; times are CUDA events on the launching stream inside the timed region.
    n_tuples = ds.num_tuples
    R_u, R_i = ds.distinct_users, ds.distinct_items
    bytes_u = n_tuples * (4 * d + 4) + 4 * (num_users + 1) + 4 * R_u * d
    bytes_v = n_tuples * (4 * d + 4) + 4 * (num_items + 1) + 4 * R_i * d + 4 * n_tuples
    # (the block-subspace models run d / block_size block sweeps per side instead of one half-step)
    t_rows = sum(stage_ms.get(k, 0.0) for k in ("step_U", "step_V", "block_U", "block_V")) * 1e-3
    peaks = measured_peaks()
    peak, peak_src = peaks["hbm_gbs"], peaks["source"]
    # TF32 dense peak: half the measured bf16 cuBLAS rate (same tensor datapath at half the K per instruction);
    # sustained figure because the row kernel runs inside a long step.
    tf32_peak = 0.5 * peaks["bf16_tflops_sustained"]
    achieved = (bytes_u + bytes_v) / world / t_rows / 1e9 if t_rows > 0 else 0.0
    fl = epoch_flops(n_tuples, R_u, R_i, num_users, num_items, d)
    row_flops = fl["syrk"] + fl["cholesky"]          # what the reference's algorithm does in the two half-steps
    useful_tf = row_flops / world / t_rows / 1e12 if t_rows > 0 else 0.0

    def stage(name, nbytes=None, flops=None):
        t = stage_ms.get(name, 0.0) * 1e-3
        if t <= 0:
            return None
        e = {"ms": t * 1e3}
        if nbytes is not None:
            e["hbm_gbs"] = nbytes / world / t / 1e9
            e["hbm_frac"] = e["hbm_gbs"] / peak
        if flops is not None:
            e["tflops"] = flops / world / t / 1e12
            e["tensor_frac_of_tf32_peak"] = e["tflops"] / tf32_peak
        return e

    stages = {
        # true dense contractions: tensor pipe (3xTF32: three MMA passes per useful flop)
        "gramian_Uz": stage("gramian_Uz", 4 * num_users * d + 4 * num_users + 4 * d * d, 2 * num_users * d * d),
        "gramian_V": stage("gramian_V", 4 * num_items * d + 4 * d * d, 2 * num_items * d * d),
        # gather-bound passes
        "user_loss": stage("user_loss", n_tuples * (4 * d + 4) + 4 * num_users * d + 4 * num_users, 2 * n_tuples * d),
        "quadform": stage("quadform", 4 * num_users * d + 4 * d * d, num_users * (2 * d * d + 2 * d)),
        "xi": stage("xi", 4 * int(num_users * cfg["sampling_ratio"] if cfg["use_snr"] else num_users) * cfg["xi_iterations"] * 2),
        "weights": stage("weights", 8 * num_users),
        "step_U": stage("step_U", bytes_u, fl["syrk"] / 2 + R_u * (d ** 3 / 3 + 2 * d * d)),
        "step_V": stage("step_V", bytes_v, fl["syrk"] / 2 + R_i * (d ** 3 / 3 + 2 * d * d)),
        "tridiag": stage("tridiag"), "rotate": stage("rotate", None, 2 * num_items * d * d),
    }
    traffic, traffic_src = profile_traffic(args, world, cfg)
    roofline = {"bound": "hbm", "kernel": "row kernels (step_U + step_V launches: CSR gather + SYRK + solve; direct d x d "
                                          "Cholesky for rows > 128 entries, dual form for the rest; fp32 accuracy "
                                          "from fp16 / tf32 hi+lo operand pairs with fp32 accumulation)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_epoch": bytes_u + bytes_v,
                # the same launches against the tensor roofline: flops of the REFERENCE's algorithm (lower-triangle SYRK
                # + one d x d Cholesky per row); the error-compensated operands (fp16 hi/lo pairs in the direct kernel's
                # SYRK, tf32 hi/lo elsewhere) issue three MMA passes per useful flop, the dual form
                # executes fewer flops than the reference's algorithm for the rows it serves
                "tensor": {"useful_tflops": useful_tf, "peak_tflops_tf32": tf32_peak,
                           "peak_source": "0.5 x bf16_tflops_sustained of MEASURED_PEAKS.json" if peaks["measured"]
                                          else "0.5 x fallback bf16 rate (B200_PROFILING.md)",
                           "frac": useful_tf / tf32_peak, "algorithmic_flops_per_epoch": row_flops},
                "bound_times_ms": {"t_hbm": (bytes_u + bytes_v) / world / (peak * 1e9) * 1e3,
                                   "t_tensor_useful": row_flops / world / (tf32_peak * 1e12) * 1e3,
                                   "t_measured": t_rows * 1e3},
                "kernel_share_of_step": t_rows * 1e3 / max(1e-9, sum(stage_ms.values())),
                "stages": {k: v for k, v in stages.items() if v},
                "stage_ms": stage_ms}
    if args.profile_stages and rank == 0:
        print(json.dumps(stage_ms, indent=1), file=sys.stderr)

    if rank != 0:
        return
    out = {"metric": f"{args.model}_epoch_row_solves_per_s", "value": value, "unit": "row-solves/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "epoch_s": ms_per_step * 1e-3,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "dtype_note": "fp32 factors and results; tensor-core products use error-compensated fp16 / tf32 hi+lo operand "
                         "pairs (three MMA passes, fp32 accumulation): parity with the fp32 reference <= 1e-4",
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
