"""torchrun script: the row-sharded multi-GPU epoch must match the CPU oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_parity.py

Every rank builds the same dataset and model, owns a contiguous block of user rows and of item
rows (balanced on history length), and exchanges factor blocks / partial Gramians / losses over
NCCL inside the library.  Rank 0 checks the result against the oracle; all ranks check that their
replicated state is bit-identical to rank 0's."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = helpers.load_pkg()
    ctx = pkg.Context(local)
    uid = [pkg.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.init_comm(rank, world, uid[0])
    nu, ni = 500, 400
    users, items = helpers.synth_tuples(nu, ni, 30, seed=5, heavy_rows=[(3, 200), (4, 129)], empty_users=(9,),
                                        empty_items=(2,))
    failures = 0
    cases = [("safer2", 32, dict(uobs_weight=0.004, reg=0.004, bandwidth=0.15, use_snr=1, sampling_ratio=0.5, snr_seed=3)),
             ("safer2", 128, dict(uobs_weight=0.004, reg=0.004, bandwidth=0.15)),
             ("safer2", 256, dict(uobs_weight=0.002, reg=0.002, bandwidth=0.18, use_snr=1, sampling_ratio=0.1, snr_seed=1)),
             ("ials", 32, dict(uobs_weight=0.1, reg=0.003)),
             ("erm_mf", 32, dict(uobs_weight=0.004, reg=0.005)),
             ("cvar_mf", 32, dict(uobs_weight=0.008, reg=0.002, stepsize=0.4)),
             # block-subspace solvers: every rank refreshes the cached predictions of the rows it is about to solve
             ("ialspp", 32, dict(uobs_weight=0.1, reg=0.003, block_size=8)),
             ("safer2pp", 32, dict(uobs_weight=0.004, reg=0.004, bandwidth=0.15, block_size=8)),
             ("safer2pp", 128, dict(uobs_weight=0.004, reg=0.004, bandwidth=0.15, block_size=64))]
    only = os.environ.get("DIST_PARITY_CASES")  # e.g. "safer2:256,ialspp:32" (box time at N = 8 is charged 8x)
    if only:
        keep = {tuple(x.split(":")) for x in only.split(",")}
        cases = [c for c in cases if (c[0], str(c[1])) in keep]
    for name, d, cfg in cases:
        ds = pkg.Dataset(ctx, users, items)
        m = pkg.Model(ctx, nu, ni, model=name, dim=d, **cfg)
        m.init_factors(11)
        m.initialize(ds)
        for _ in range(2):
            m.train(ds)
        U, V = m.factors()
        st = m.state()
        # replicated state must be identical on every rank
        blob = np.concatenate([U.ravel(), V.ravel(), st["z"], st["loss"], [st["xi"]]]).astype(np.float32)
        t = torch.from_numpy(blob).cuda()
        ref = t.clone()
        dist.broadcast(ref, src=0)
        same = bool(torch.equal(t, ref))
        if rank == 0:
            from oracle import loader as O
            ods = O.Dataset.from_tuples(users, items)
            om = O.Model(nu, ni, init_seed=11, model=name, dim=d, **cfg)
            om.initialize(ods)
            for _ in range(2):
                om.train(ods)
            Uo, Vo = om.factors()
            so = om.state()
            eu, ev = helpers.rel_fro(U, Uo), helpers.rel_fro(V, Vo)
            ok = eu < 5e-4 and ev < 5e-4 and abs(st["xi"] - so["xi"]) < 1e-3 and np.allclose(st["loss"], so["loss"], rtol=5e-4, atol=1e-6)
            print(f"[dist_parity] world={world} {name} d={d}: relF U={eu:.2e} V={ev:.2e} xi {st['xi']:.6f}/{so['xi']:.6f} {'OK' if ok else 'FAIL'}", flush=True)
            failures += 0 if ok else 1
        if not same:
            print(f"[dist_parity] rank {rank}: replicated state differs from rank 0 for {name} d={d}", flush=True)
            failures += 1
        # row-sharded host legs: upload own rows + all-gather == full upload; download writes own rows only
        rngs = np.random.default_rng(5)
        Uh = rngs.standard_normal((nu, d)).astype(np.float32)
        Vh = rngs.standard_normal((ni, d)).astype(np.float32)
        m.upload_factors_sharded(ds, Uh, Vh)
        ctx.sync()
        U2, V2 = m.factors()
        ub = pkg.partition_rows(ds.csr(0, ds.max_user + 1)[0], world)
        ib = pkg.partition_rows(ds.csr(1, ds.max_item + 1)[0], world)
        Ud = np.full((nu, d), -7.0, np.float32)
        Vd = np.full((ni, d), -7.0, np.float32)
        m.factors_sharded(ds, Ud, Vd)
        own_ok = (np.array_equal(Ud[ub[rank]:ub[rank + 1]], Uh[ub[rank]:ub[rank + 1]]) and
                  np.array_equal(Vd[ib[rank]:ib[rank + 1]], Vh[ib[rank]:ib[rank + 1]]) and
                  np.all(Ud[:ub[rank]] == -7.0) and np.all(Ud[ub[rank + 1]:ds.max_user + 1] == -7.0))
        # train_to_host: the overlapped device->host copies must deliver this rank's rows of the NEW factors
        Ut = np.full((nu, d), -7.0, np.float32)
        Vt = np.full((ni, d), -7.0, np.float32)
        m.train_to_host(ds, Ut, Vt)
        U3, V3 = m.factors()
        own_ok = (own_ok and np.array_equal(Ut[ub[rank]:ub[rank + 1]], U3[ub[rank]:ub[rank + 1]]) and
                  np.array_equal(Vt[ib[rank]:ib[rank + 1]], V3[ib[rank]:ib[rank + 1]]) and
                  np.all(Ut[:ub[rank]] == -7.0) and np.all(Vt[:ib[rank]] == -7.0))
        if not (np.array_equal(U2, Uh) and np.array_equal(V2, Vh) and own_ok):
            print(f"[dist_parity] rank {rank}: sharded factor transfer mismatch for {name} d={d}", flush=True)
            failures += 1
        m.close()
        ds.close()
    f = torch.tensor([failures], device="cuda")
    dist.all_reduce(f)
    ctx.close()
    dist.destroy_process_group()
    if int(f.item()):
        sys.exit(1)
    if rank == 0:
        print("[dist_parity] PASS", flush=True)


if __name__ == "__main__":
    main()
