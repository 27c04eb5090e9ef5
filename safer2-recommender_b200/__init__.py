"""frecsys_b200 — B200-native ALS hot path of frecsys (riktor/safer2-recommender).

This package is a thin ctypes binding over the C ABI declared in
``include/frecsys_b200.h`` and implemented by hand-written sm_100a kernels in
``csrc/`` (built in-tree as ``libfrecsys_b200.so``).  The C++ host-side mirror of
the reference's ``frecsys::Recommender`` classes lives in ``include/frecsys``.

There is no CPU fallback: loading fails loudly if the CUDA library is missing,
and every compute call fails if no CUDA device is present.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# FRECSYS_B200_LIB: diagnostic override (experimental builds of the same library)
LIB_PATH = os.environ.get("FRECSYS_B200_LIB") or os.path.join(HERE, "libfrecsys_b200.so")

MODEL_IDS = {"ials": 0, "ialspp": 1, "erm_mf": 2, "cvar_mf": 3, "safer2": 4, "safer2pp": 5}

EXPORTS = [
    "frx_last_error", "frx_context_create", "frx_context_destroy", "frx_context_sync",
    "frx_context_stream", "frx_comm_unique_id", "frx_context_init_comm", "frx_partition_rows", "frx_dataset_create", "frx_model_upload_factors_sharded", "frx_model_get_factors_sharded", "frx_model_train_to_host", "frx_model_save", "frx_model_load",
    "frx_dataset_destroy", "frx_dataset_info", "frx_dataset_get_csr", "frx_model_create",
    "frx_model_destroy", "frx_model_init_factors", "frx_model_set_factors", "frx_model_get_factors", "frx_model_upload_factors",
    "frx_model_initialize", "frx_model_train", "frx_model_stage", "frx_model_get_state",
    "frx_model_set_state", "frx_model_compute_stats", "frx_model_last_snr", "frx_model_evaluate",
    "frx_context_launch_count", "frx_context_set_profiling", "frx_context_stage_times", "frx_gramian", "frx_sym_tridiag", "frx_model_set_residual_stats", "frx_model_get_residuals",
]


class FrxConfig(C.Structure):
    """Mirror of ``frx_config`` (include/frecsys_b200.h)."""
    _fields_ = [
        ("model", C.c_int), ("dim", C.c_int),
        ("reg", C.c_float), ("reg_exp", C.c_float), ("uobs_weight", C.c_float),
        ("stdev", C.c_float), ("alpha", C.c_float), ("bandwidth", C.c_float),
        ("stepsize", C.c_float),
        ("xi_iterations", C.c_int), ("pd_iterations", C.c_int),
        ("use_epanechnikov", C.c_int), ("use_snr", C.c_int),
        ("sampling_ratio", C.c_float), ("use_cg", C.c_int), ("cg_tol", C.c_float),
        ("cg_max_it", C.c_int), ("block_size", C.c_int), ("snr_seed", C.c_uint),
    ]


# run_model.cc:128-230 defaults
DEFAULTS = dict(model="ials", dim=8, reg=0.002, reg_exp=1.0, uobs_weight=0.1, stdev=0.1,
                alpha=0.3, bandwidth=1.0, stepsize=0.1, xi_iterations=5, pd_iterations=1,
                use_epanechnikov=0, use_snr=0, sampling_ratio=0.1, use_cg=0, cg_tol=1e-10,
                cg_max_it=100, block_size=64, snr_seed=0)


def make_config(**kw):
    d = dict(DEFAULTS)
    d.update(kw)
    c = FrxConfig()
    for k, v in d.items():
        if k == "model":
            v = MODEL_IDS[v.lower()] if isinstance(v, str) else int(v)
        setattr(c, k, v)
    return c


class FrxError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libfrecsys_b200.so; raises if the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FrxError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C safer2-recommender_b200/csrc). frecsys_b200 has no CPU fallback.")
    try:
        import torch  # noqa: F401  (loads the NCCL shared library torch bundles before ours resolves libnccl.so.2)
    except Exception:
        pass
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, ip, fp, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_double)
    L.frx_last_error.restype = C.c_char_p
    L.frx_context_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    L.frx_context_destroy.argtypes = [vp]
    L.frx_context_destroy.restype = None
    L.frx_context_sync.argtypes = [vp]
    L.frx_context_stream.argtypes = [vp]
    L.frx_context_stream.restype = vp
    L.frx_comm_unique_id.argtypes = [vp]
    L.frx_context_init_comm.argtypes = [vp, C.c_int, C.c_int, vp]
    L.frx_partition_rows.argtypes = [ip, C.c_int, C.c_int, C.c_int, ip]
    L.frx_dataset_create.argtypes = [vp, C.c_int, ip, ip, C.POINTER(vp)]
    L.frx_dataset_destroy.argtypes = [vp]
    L.frx_dataset_destroy.restype = None
    L.frx_dataset_info.argtypes = [vp, ip]
    L.frx_dataset_get_csr.argtypes = [vp, C.c_int, C.c_int, ip, ip, ip]
    L.frx_model_create.argtypes = [vp, C.POINTER(FrxConfig), C.c_int, C.c_int, C.POINTER(vp)]
    L.frx_model_destroy.argtypes = [vp]
    L.frx_model_destroy.restype = None
    L.frx_model_init_factors.argtypes = [vp, C.c_uint]
    L.frx_model_set_factors.argtypes = [vp, fp, fp]
    L.frx_model_get_factors.argtypes = [vp, fp, fp]
    L.frx_model_upload_factors.argtypes = [vp, fp, fp]
    L.frx_model_upload_factors_sharded.argtypes = [vp, vp, fp, fp]
    L.frx_model_get_factors_sharded.argtypes = [vp, vp, fp, fp]
    L.frx_model_train_to_host.argtypes = [vp, vp, fp, fp]
    L.frx_model_save.argtypes = [vp, C.c_char_p]
    L.frx_model_load.argtypes = [vp, C.c_char_p]
    L.frx_model_initialize.argtypes = [vp, vp]
    L.frx_model_train.argtypes = [vp, vp]
    L.frx_model_stage.argtypes = [vp, vp, C.c_int]
    L.frx_model_get_state.argtypes = [vp, fp, fp, fp, fp, fp, fp]
    L.frx_model_set_state.argtypes = [vp, fp, fp, C.c_float]
    L.frx_model_compute_stats.argtypes = [vp, vp, dp]
    L.frx_model_last_snr.argtypes = [vp, ip, ip, ip]
    L.frx_model_evaluate.argtypes = [vp, vp, vp, ip, C.c_int, ip, fp, fp, ip, fp]
    L.frx_context_launch_count.argtypes = [vp]
    L.frx_context_launch_count.restype = C.c_longlong
    L.frx_context_set_profiling.argtypes = [vp, C.c_int]
    L.frx_context_stage_times.argtypes = [vp, C.c_char_p, C.c_int, fp, C.c_int]
    L.frx_gramian.argtypes = [vp, fp, C.c_int, C.c_int, fp, fp]
    L.frx_sym_tridiag.argtypes = [vp, fp, C.c_int, fp, fp, fp]
    L.frx_model_set_residual_stats.argtypes = [vp, C.c_int]
    L.frx_model_get_residuals.argtypes = [vp, fp, C.c_int]
    _lib = L
    return L


def _check(rc):
    if rc < 0:
        raise FrxError(f"frecsys_b200 error {rc}: {lib().frx_last_error().decode()}")
    return rc


def _fp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int))


class Context:
    """One GPU (frx_context).  ``stream`` may be a raw cudaStream_t (int), e.g.
    ``torch.cuda.current_stream().cuda_stream``, so torch CUDA events time our launches."""

    def __init__(self, device=0, stream=None):
        self.h = C.c_void_p()
        _check(lib().frx_context_create(device, C.c_void_p(stream) if stream else None, C.byref(self.h)))
        self.device = device

    def sync(self):
        _check(lib().frx_context_sync(self.h))

    def launch_count(self):
        return int(lib().frx_context_launch_count(self.h))

    def set_profiling(self, on):
        _check(lib().frx_context_set_profiling(self.h, int(on)))

    def stage_times(self):
        buf = C.create_string_buffer(8192)
        ms = np.zeros(512, np.float32)
        n = _check(lib().frx_context_stage_times(self.h, buf, len(buf), _fp(ms), len(ms)))
        names = buf.value.decode().split(";") if n else []
        return list(zip(names, [float(x) for x in ms[:n]]))

    def init_comm(self, rank, world, unique_id):
        _check(lib().frx_context_init_comm(self.h, rank, world, unique_id))

    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(128)
        _check(lib().frx_comm_unique_id(buf))
        return buf.raw

    def gramian(self, E, w=None):
        E = np.ascontiguousarray(E, np.float32)
        w = None if w is None else np.ascontiguousarray(w, np.float32)
        out = np.zeros((E.shape[1], E.shape[1]), np.float32)
        _check(lib().frx_gramian(self.h, _fp(E), E.shape[0], E.shape[1], _fp(w), _fp(out)))
        return out

    def sym_tridiag(self, G):
        """(H, tdiag, tsub) with G = H T H^T, T tridiagonal (tsub[j] = T[j][j-1]); d = 128 or 256 (the kernel of
        the dual-form row path)."""
        G = np.ascontiguousarray(G, np.float32)
        d = G.shape[0]
        H = np.zeros((d, d), np.float32)
        td = np.zeros(d, np.float32)
        ts = np.zeros(d, np.float32)
        _check(lib().frx_sym_tridiag(self.h, _fp(G), d, _fp(H), _fp(td), _fp(ts)))
        return H, td, ts

    def close(self):
        if self.h:
            lib().frx_context_destroy(self.h)
            self.h = None


def partition_rows(ptr, world, row_unit=-1):
    """Row ranges per rank for the row-sharded epoch (host-only; mirrors what frx_dataset_create uses)."""
    ptr = np.ascontiguousarray(ptr, np.int32)
    out = np.zeros(world + 1, np.int32)
    _check(lib().frx_partition_rows(_ip(ptr), len(ptr) - 1, world, row_unit, _ip(out)))
    return out


def read_csv_tuples(path):
    """(users, items) int32 arrays in file order; `uid,sid` with a header line (dataset.h:71-99)."""
    arr = np.loadtxt(path, delimiter=",", skiprows=1, dtype=np.int64, ndmin=2)
    return arr[:, 0].astype(np.int32), arr[:, 1].astype(np.int32)


class Dataset:
    """Device-resident interactions (frx_dataset) built from the tuple list in file order."""

    def __init__(self, ctx, users, items):
        self.ctx = ctx
        users = np.ascontiguousarray(users, np.int32)
        items = np.ascontiguousarray(items, np.int32)
        assert users.shape == items.shape
        self.h = C.c_void_p()
        _check(lib().frx_dataset_create(ctx.h, len(users), _ip(users), _ip(items), C.byref(self.h)))
        info = (C.c_int * 5)()
        _check(lib().frx_dataset_info(self.h, info))
        self.max_user, self.max_item, self.num_tuples, self.distinct_users, self.distinct_items = list(info)

    @classmethod
    def from_csv(cls, ctx, path):
        u, i = read_csv_tuples(path)
        return cls(ctx, u, i)

    def csr(self, by_item, nrows):
        ptr = np.zeros(nrows + 1, np.int32)
        ids = np.zeros(max(1, self.num_tuples), np.int32)
        tup = np.zeros(max(1, self.num_tuples), np.int32)
        _check(lib().frx_dataset_get_csr(self.h, int(by_item), nrows, _ip(ptr), _ip(ids), _ip(tup)))
        return ptr, ids[:self.num_tuples], tup[:self.num_tuples]

    def close(self):
        if self.h:
            lib().frx_dataset_destroy(self.h)
            self.h = None


class Model:
    """One recommender (frx_model): factors + per-user state on the device."""

    def __init__(self, ctx, num_users, num_items, **cfg):
        self.ctx = ctx
        self.cfg = make_config(**cfg)
        self.num_users, self.num_items, self.dim = num_users, num_items, self.cfg.dim
        self.h = C.c_void_p()
        _check(lib().frx_model_create(ctx.h, C.byref(self.cfg), num_users, num_items, C.byref(self.h)))

    def init_factors(self, seed):
        _check(lib().frx_model_init_factors(self.h, seed))

    def set_factors(self, U, V):
        U = None if U is None else np.ascontiguousarray(U, np.float32)
        V = None if V is None else np.ascontiguousarray(V, np.float32)
        _check(lib().frx_model_set_factors(self.h, _fp(U), _fp(V)))
        self.ctx.sync()

    def upload_factors(self, U, V):
        """Overwrite factors only (state untouched), asynchronously; U, V must stay alive
        (ideally pinned) until the context is synchronised."""
        _check(lib().frx_model_upload_factors(self.h, _fp(U), _fp(V)))

    def upload_factors_sharded(self, ds, U, V):
        """Multi-rank H2D leg: this rank uploads only the rows it owns, then the blocks are all-gathered."""
        _check(lib().frx_model_upload_factors_sharded(self.h, ds.h, _fp(U), _fp(V)))

    def factors_sharded(self, ds, U, V):
        """Multi-rank D2H leg: only this rank's rows of U and V are written into the host arrays."""
        _check(lib().frx_model_get_factors_sharded(self.h, ds.h, _fp(U), _fp(V)))
        return U, V

    def factors(self, U=None, V=None):
        U = np.zeros((self.num_users, self.dim), np.float32) if U is None else U
        V = np.zeros((self.num_items, self.dim), np.float32) if V is None else V
        _check(lib().frx_model_get_factors(self.h, _fp(U), _fp(V)))
        return U, V

    def initialize(self, ds):
        _check(lib().frx_model_initialize(self.h, ds.h))

    def save(self, path):
        _check(lib().frx_model_save(self.h, os.fsencode(path)))

    def load(self, path):
        _check(lib().frx_model_load(self.h, os.fsencode(path)))

    def train_to_host(self, ds, U, V):
        """One epoch, then U and V (this rank's rows) in the host arrays; the copy of U overlaps the item half-step."""
        _check(lib().frx_model_train_to_host(self.h, ds.h, _fp(U), _fp(V)))
        return U, V

    def train(self, ds):
        _check(lib().frx_model_train(self.h, ds.h))

    def stage(self, ds, stage):
        _check(lib().frx_model_stage(self.h, ds.h, stage))

    def state(self):
        z = np.zeros(self.num_users, np.float32)
        loss = np.zeros(self.num_users, np.float32)
        hs = np.zeros(self.num_users, np.float32)
        ireg = np.zeros(self.num_items, np.float32)
        sc = np.zeros(3, np.float32)
        G = np.zeros((self.dim, self.dim), np.float32)
        _check(lib().frx_model_get_state(self.h, _fp(z), _fp(loss), _fp(hs), _fp(ireg), _fp(sc), _fp(G)))
        return dict(z=z, loss=loss, hist_size=hs, item_reg=ireg, xi=float(sc[0]),
                    weighted_loss=float(sc[1]), mean_weight=float(sc[2]), gramian=G)

    def scalars(self):
        sc = np.zeros(3, np.float32)
        _check(lib().frx_model_get_state(self.h, None, None, None, None, _fp(sc), None))
        return dict(xi=float(sc[0]), weighted_loss=float(sc[1]), mean_weight=float(sc[2]))

    def set_state(self, z=None, loss=None, xi=0.0):
        z = None if z is None else np.ascontiguousarray(z, np.float32)
        loss = None if loss is None else np.ascontiguousarray(loss, np.float32)
        _check(lib().frx_model_set_state(self.h, _fp(z), _fp(loss), float(xi)))

    def stats(self, ds):
        out = (C.c_double * 6)()
        _check(lib().frx_model_compute_stats(self.h, ds.h, out))
        keys = ["loss", "loss_observed", "loss_unobserved", "loss_reg", "loss_reg_user", "loss_reg_item"]
        return dict(zip(keys, list(out)))

    def set_residual_stats(self, on=True):
        _check(lib().frx_model_set_residual_stats(self.h, 1 if on else 0))

    def residuals(self):
        """[iterations x 3] U / V / z residual norms of the last train() (--print_residual_stats)."""
        out = np.zeros(3 * 64, np.float32)
        n = _check(lib().frx_model_get_residuals(self.h, _fp(out), 64))
        return out[:3 * n].reshape(n, 3)

    def last_snr(self):
        ni, ns = C.c_int(), C.c_int()
        _check(lib().frx_model_last_snr(self.h, C.byref(ni), C.byref(ns), None))
        out = np.zeros((ni.value, ns.value), np.int32)
        if out.size:
            _check(lib().frx_model_last_snr(self.h, C.byref(ni), C.byref(ns), _ip(out)))
        return out

    def evaluate(self, tr, te, k_list=(5, 10, 20, 50, 100), want_topk=False, want_folded=False):
        ks = np.asarray(k_list, np.int32)
        nu = _check(lib().frx_model_evaluate(self.h, tr.h, te.h, _ip(ks), len(ks), None, None, None, None, None))
        ids = np.zeros(nu, np.int32)
        rec = np.zeros((nu, len(ks)), np.float32)
        ndcg = np.zeros((nu, len(ks)), np.float32)
        topk = np.zeros((nu, int(ks.max())), np.int32) if want_topk else None
        folded = np.zeros((nu, self.dim), np.float32) if want_folded else None
        _check(lib().frx_model_evaluate(self.h, tr.h, te.h, _ip(ks), len(ks), _ip(ids), _fp(rec), _fp(ndcg),
                                        _ip(topk), _fp(folded)))
        return dict(user_ids=ids, recall=rec, ndcg=ndcg, topk=topk, folded=folded)

    def close(self):
        if self.h:
            lib().frx_model_destroy(self.h)
            self.h = None
