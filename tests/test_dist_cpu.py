"""CPU (gloo, world_size 2) test of the row-sharded epoch's host-side logic: the product's partition
function (frx_partition_rows) decides which rank owns which rows; every rank solves only its rows with
the oracle and the factor blocks / partial Gramians / losses are exchanged with torch.distributed
collectives at exactly the points where the CUDA library calls NCCL (csrc/frx_api.cu: all-gather after
each half-step, all-reduce of the partial Gramians, all-gather of the losses).  The result must equal the
single-process oracle epoch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers

CFG = dict(model="safer2", dim=16, uobs_weight=0.004, reg=0.004, bandwidth=0.15, use_snr=1, sampling_ratio=0.5, snr_seed=4)
NU, NI = 240, 180


def _data():
    return helpers.synth_tuples(NU, NI, 20, seed=17, heavy_rows=[(3, 150)], empty_users=(5,), empty_items=(9,))


def _allsum(a):
    t = torch.from_numpy(np.ascontiguousarray(a))
    dist.all_reduce(t)
    return t.numpy()


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, helpers.ROOT)
    from oracle import loader as O
    pkg = helpers.load_pkg()
    users, items = _data()
    ds = O.Dataset.from_tuples(users, items)
    uptr = ds.csr(0, NU)[0]
    iptr = ds.csr(1, NI)[0]
    ub = pkg.partition_rows(uptr, world)   # the product's own partition (host-only entry point)
    ib = pkg.partition_rows(iptr, world)
    m = O.Model(NU, NI, init_seed=21, **CFG)
    m.initialize(ds)                       # replicated (the library shards the loss pass; same values)
    umask = np.zeros((NU, 1), np.float32)
    umask[ub[rank]:ub[rank + 1]] = 1
    vmask = np.zeros((NI, 1), np.float32)
    vmask[ib[rank]:ib[rank + 1]] = 1
    assert np.all(_allsum(umask.copy()) == 1) and np.all(_allsum(vmask.copy()) == 1)  # every row has one owner
    for _ in range(2):
        m.stage(ds, 0)                                         # z: replicated
        # --- user half-step on own rows, then all-gather of the U blocks
        m.set_range(ub[rank], ub[rank + 1])
        m.stage(ds, 1)
        U = _allsum(m.factors()[0] * umask)
        m.put_factors(U=U)
        # --- partial U^T diag(z) U over an even share of the rows, all-reduce
        z = m.state()["z"]
        b, e = NU * rank // world, NU * (rank + 1) // world
        m.set_gz_override(_allsum(O.gramian(U[b:e], z[b:e])))
        # --- item half-step on own rows, all-gather of the V blocks
        m.set_range(ib[rank], ib[rank + 1])
        m.stage(ds, 2)
        V = _allsum(m.factors()[1] * vmask)
        m.put_factors(V=V)
        # --- partial V^T V, all-reduce
        b, e = NI * rank // world, NI * (rank + 1) // world
        m.set_item_gramian(_allsum(O.gramian(V[b:e])))
        # --- per-user loss on own rows, all-gather
        m.set_range(ub[rank], ub[rank + 1])
        m.stage(ds, 4)
        st = m.state()
        loss = _allsum(st["loss"] * umask[:, 0])
        m.set_state(z=st["z"], loss=loss, xi=st["xi"])
        m.set_range(0, 2 ** 30)
        m.stage(ds, 5)                                         # xi: replicated, identical seeds
    U, V = m.factors()
    st = m.state()
    if rank == 0:
        np.savez(out_path, U=U, V=V, xi=st["xi"], loss=st["loss"], ub=ub, ib=ib)
    dist.destroy_process_group()


def test_partition_rows_covers_and_balances():
    pkg = helpers.load_pkg()
    rng = np.random.default_rng(0)
    n = rng.integers(0, 200, 1000)
    ptr = np.concatenate([[0], np.cumsum(n)]).astype(np.int32)
    for world in (1, 2, 3, 8):
        rb = pkg.partition_rows(ptr, world, row_unit=256)
        assert rb[0] == 0 and rb[-1] == 1000 and np.all(np.diff(rb) >= 0)
        cost = np.array([ptr[rb[k + 1]] - ptr[rb[k]] + 256 * np.count_nonzero(n[rb[k]:rb[k + 1]]) for k in range(world)])
        assert cost.max() <= cost.sum() / world + 200 + 256
        # built-in cost model (what frx_dataset_create uses): rows of at most 128 entries go to the dual-form
        # kernel at a flat cost, longer rows to the direct kernel at length + 480
        rb = pkg.partition_rows(ptr, world)
        assert rb[0] == 0 and rb[-1] == 1000 and np.all(np.diff(rb) >= 0)
        c = np.where(n == 0, 0, np.where(n <= 128, 150, n + 480))
        cost = np.array([c[rb[k]:rb[k + 1]].sum() for k in range(world)])
        assert cost.max() <= cost.sum() / world + c.max()
    # degenerate: fewer rows than ranks, empty matrix
    assert pkg.partition_rows(np.array([0, 5, 9], np.int32), 8)[-1] == 2
    assert list(pkg.partition_rows(np.array([0], np.int32), 4)) == [0, 0, 0, 0, 0]


def test_row_sharded_epoch_world2_gloo(tmp_path):
    from oracle import loader as O
    out = str(tmp_path / "r0.npz")
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    users, items = _data()
    ds = O.Dataset.from_tuples(users, items)
    m = O.Model(NU, NI, init_seed=21, **CFG)
    m.initialize(ds)
    for _ in range(2):
        m.train(ds)
    U, V = m.factors()
    st = m.state()
    assert helpers.rel_fro(got["U"], U) < 1e-5
    assert helpers.rel_fro(got["V"], V) < 1e-5
    assert abs(float(got["xi"]) - st["xi"]) < 1e-5
    np.testing.assert_allclose(got["loss"], st["loss"], rtol=1e-4, atol=1e-7)
    assert 0 < got["ub"][1] < NU and 0 < got["ib"][1] < NI


PP_CFG = dict(dim=16, block_size=8, uobs_weight=0.05, reg=0.01)


def _pp_worker(rank, world, port, out_path, name):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, helpers.ROOT)
    from oracle import loader as O
    pkg = helpers.load_pkg()
    users, items = _data()
    ds = O.Dataset.from_tuples(users, items)
    ub = pkg.partition_rows(ds.csr(0, NU)[0], world)
    ib = pkg.partition_rows(ds.csr(1, NI)[0], world)
    cfg = dict(PP_CFG, model=name)
    if name == "safer2pp":
        cfg.update(uobs_weight=0.004, reg=0.004, bandwidth=0.15)
    m = O.Model(NU, NI, init_seed=21, **cfg)
    m.initialize(ds)
    umask = np.zeros((NU, 1), np.float32)
    umask[ub[rank]:ub[rank + 1]] = 1
    vmask = np.zeros((NI, 1), np.float32)
    vmask[ib[rank]:ib[rank + 1]] = 1
    d, B = cfg["dim"], cfg["block_size"]
    pred = np.zeros(len(users), np.float32)       # replicated in memory, current only for the rows about to be solved
    for _ in range(2):
        if name == "safer2pp":
            m.set_range(0, 2 ** 30)
            m.stage(ds, 0)                                      # z (all users, safer2pp.h:847-856): replicated
        for start in range(0, d, B):
            end = min(start + B, d)
            # csrc/frx_api.cu stage_block: refresh the cached predictions of this rank's rows, block sweep, all-gather
            m.set_range(ub[rank], ub[rank + 1])
            m.predict_rows(ds, 0, pred)
            m.block_step(ds, 0, start, end, pred)
            m.put_factors(U=_allsum(m.factors()[0] * umask))
            m.set_range(ib[rank], ib[rank + 1])
            m.predict_rows(ds, 1, pred)
            m.block_step(ds, 1, start, end, pred)
            m.put_factors(V=_allsum(m.factors()[1] * vmask))
        if name == "safer2pp":
            # item Gramian, per-user loss on own rows (from factors here: the product refreshes its cache and reads
            # the same dot products), all-gather, xi replicated
            V = m.factors()[1]
            b, e = NI * rank // world, NI * (rank + 1) // world
            m.set_item_gramian(_allsum(O.gramian(V[b:e])))
            m.set_range(ub[rank], ub[rank + 1])
            m.stage(ds, 4)
            st = m.state()
            m.set_state(z=st["z"], loss=_allsum(st["loss"] * umask[:, 0]), xi=st["xi"])
            m.set_range(0, 2 ** 30)
            m.stage(ds, 5)
    U, V = m.factors()
    if rank == 0:
        np.savez(out_path, U=U, V=V, xi=m.state()["xi"])
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["ialspp", "safer2pp"])
def test_block_solvers_row_sharded_world2_gloo(tmp_path, name):
    """The multi-GPU protocol of the block-subspace models (csrc/frx_api.cu stage_block with several ranks): the
    prediction cache is replicated in memory but only kept current for the rows a rank is about to solve (recomputed
    from the all-gathered factors), the updated block is all-gathered after every sweep.  Two oracle processes
    following that protocol over gloo must reproduce the single-process epoch (whose cache is incremental,
    ialspp.h:136-143) to fp32 rounding."""
    from oracle import loader as O
    out = str(tmp_path / "pp.npz")
    port = 29900 + os.getpid() % 90
    mp.spawn(_pp_worker, args=(2, port, out, name), nprocs=2, join=True)
    got = np.load(out)
    users, items = _data()
    ds = O.Dataset.from_tuples(users, items)
    cfg = dict(PP_CFG, model=name)
    if name == "safer2pp":
        cfg.update(uobs_weight=0.004, reg=0.004, bandwidth=0.15)
    m = O.Model(NU, NI, init_seed=21, **cfg)
    m.initialize(ds)
    for _ in range(2):
        m.train(ds)
    U, V = m.factors()
    assert helpers.rel_fro(got["U"], U) < 2e-5, helpers.rel_fro(got["U"], U)
    assert helpers.rel_fro(got["V"], V) < 2e-5, helpers.rel_fro(got["V"], V)
    if name == "safer2pp":
        assert abs(float(got["xi"]) - m.state()["xi"]) < 1e-4
