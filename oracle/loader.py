"""ctypes binding of the CPU oracle (oracle/libfrecsys_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by the product package.
"""
import ctypes as C
import hashlib
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# FRECSYS_ORACLE_FAST=1 (set by bench.py's CPU legs BEFORE this module is imported) selects the timing-only
# build: SYRK / Cholesky register-blocked and vectorised like Eigen's, different summation order.  The parity
# tests never set it.
FAST = os.environ.get("FRECSYS_ORACLE_FAST", "0") == "1"
LIB_NAME = "libfrecsys_oracle_fast.so" if FAST else "libfrecsys_oracle.so"
LIB = os.path.join(HERE, LIB_NAME)
STAMP = os.path.join(HERE, ".oracle_build_stamp_fast" if FAST else ".oracle_build_stamp")

MODEL_IDS = {"ials": 0, "ialspp": 1, "erm_mf": 2, "cvar_mf": 3, "safer2": 4, "safer2pp": 5}


class OrcConfig(C.Structure):
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


DEFAULTS = dict(model="ials", dim=8, reg=0.002, reg_exp=1.0, uobs_weight=0.1, stdev=0.1,
                alpha=0.3, bandwidth=1.0, stepsize=0.1, xi_iterations=5, pd_iterations=1,
                use_epanechnikov=0, use_snr=0, sampling_ratio=0.1, use_cg=0, cg_tol=1e-10,
                cg_max_it=100, block_size=64, snr_seed=0)


def make_config(cls=OrcConfig, **kw):
    d = dict(DEFAULTS)
    d.update(kw)
    c = cls()
    for k, v in d.items():
        if k == "model":
            v = MODEL_IDS[v] if isinstance(v, str) else int(v)
        setattr(c, k, v)
    return c


def _cpu_sig():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return hashlib.sha1(line.encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def build(force=False):
    """Compile the oracle with the reference's flags (-O3 -march=native).  The
    .so is rebuilt when the host CPU differs from the one it was built on, so a
    library built in the dev container is never run with foreign -march code."""
    import fcntl
    sig = _cpu_sig()
    with open(os.path.join(HERE, ".oracle_build_lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)  # several ranks / test workers may get here at once
        return _build_locked(sig, force)


def _build_locked(sig, force):
    srcs = [os.path.join(HERE, f) for f in ("oracle_capi.cc", "oracle_rng.cc", "frecsys_oracle.hpp", "Makefile")]
    newest = max(os.path.getmtime(s) for s in srcs)
    ok = (os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == sig
          and os.path.getmtime(LIB) >= newest)
    if ok and not force:
        return LIB
    subprocess.run(["make", "-C", HERE, "-B", LIB_NAME], check=True,
                   stdout=subprocess.DEVNULL)
    with open(STAMP, "w") as f:
        f.write(sig)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        vp, ip, fp, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_double)
        L.orc_dataset_from_csv.restype = vp
        L.orc_dataset_from_csv.argtypes = [C.c_char_p]
        L.orc_dataset_from_tuples.restype = vp
        L.orc_dataset_from_tuples.argtypes = [ip, ip, C.c_int]
        L.orc_dataset_free.argtypes = [vp]
        L.orc_dataset_info.argtypes = [vp, ip]
        L.orc_dataset_csr.argtypes = [vp, C.c_int, C.c_int, ip, ip, ip]
        L.orc_dataset_tuples.argtypes = [vp, ip, ip]
        L.orc_model_create.restype = vp
        L.orc_model_create.argtypes = [C.POINTER(OrcConfig), C.c_int, C.c_int, C.c_uint]
        L.orc_model_free.argtypes = [vp]
        L.orc_model_set_factors.argtypes = [vp, fp, fp]
        L.orc_model_get_factors.argtypes = [vp, fp, fp]
        L.orc_model_initialize.argtypes = [vp, vp]
        L.orc_model_train.argtypes = [vp, vp]
        L.orc_model_set_print_train_stats.argtypes = [vp, C.c_int]
        L.orc_model_get_state.argtypes = [vp, fp, fp, fp, fp, fp, fp]
        L.orc_model_set_state.argtypes = [vp, fp, fp, C.c_float]
        L.orc_model_get_stats.argtypes = [vp, dp]
        L.orc_model_compute_stats.argtypes = [vp, vp, dp]
        L.orc_model_last_snr.argtypes = [vp, ip, ip, ip]
        L.orc_model_stage.argtypes = [vp, vp, C.c_int]
        L.orc_model_evaluate.restype = C.c_int
        L.orc_model_evaluate.argtypes = [vp, vp, vp, ip, C.c_int, ip, fp, fp, ip, fp]
        L.orc_metric_cvar.argtypes = [fp, C.c_int, fp, C.c_int, fp]
        L.orc_gramian.argtypes = [fp, C.c_int, C.c_int, fp, fp]
        L.orc_init_factors.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_uint, fp, fp]
        L.orc_num_threads.restype = C.c_int
        L.orc_model_set_range.argtypes = [vp, C.c_int, C.c_int]
        L.orc_model_put_factors.argtypes = [vp, fp, fp]
        L.orc_model_block_step.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, fp]
        L.orc_model_predict_rows.argtypes = [vp, vp, C.c_int, fp]
        L.orc_model_set_item_gramian.argtypes = [vp, fp]
        L.orc_model_set_gz_override.argtypes = [vp, fp]
        _lib = L
    return _lib


def _fp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int))


class Dataset:
    def __init__(self, handle):
        self.h = handle
        info = (C.c_int * 5)()
        lib().orc_dataset_info(self.h, info)
        self.max_user, self.max_item, self.num_tuples, self.distinct_users, self.distinct_items = list(info)

    @classmethod
    def from_csv(cls, path):
        return cls(lib().orc_dataset_from_csv(path.encode()))

    @classmethod
    def from_tuples(cls, users, items):
        users = np.ascontiguousarray(users, dtype=np.int32)
        items = np.ascontiguousarray(items, dtype=np.int32)
        return cls(lib().orc_dataset_from_tuples(_ip(users), _ip(items), len(users)))

    def csr(self, by_item, nrows):
        ptr = np.zeros(nrows + 1, np.int32)
        ids = np.zeros(self.num_tuples, np.int32)
        tup = np.zeros(self.num_tuples, np.int32)
        lib().orc_dataset_csr(self.h, int(by_item), nrows, _ip(ptr), _ip(ids), _ip(tup))
        return ptr, ids, tup

    def tuples(self):
        u = np.zeros(self.num_tuples, np.int32)
        i = np.zeros(self.num_tuples, np.int32)
        lib().orc_dataset_tuples(self.h, _ip(u), _ip(i))
        return u, i

    def __del__(self):
        try:
            lib().orc_dataset_free(self.h)
        except Exception:
            pass


class Model:
    def __init__(self, num_users, num_items, init_seed=12345, **cfg):
        self.cfg = make_config(**cfg)
        self.num_users, self.num_items, self.dim = num_users, num_items, self.cfg.dim
        self.h = lib().orc_model_create(C.byref(self.cfg), num_users, num_items, init_seed)

    def __del__(self):
        try:
            lib().orc_model_free(self.h)
        except Exception:
            pass

    def set_factors(self, U, V):
        U = np.ascontiguousarray(U, np.float32)
        V = np.ascontiguousarray(V, np.float32)
        lib().orc_model_set_factors(self.h, _fp(U), _fp(V))

    def factors(self):
        U = np.zeros((self.num_users, self.dim), np.float32)
        V = np.zeros((self.num_items, self.dim), np.float32)
        lib().orc_model_get_factors(self.h, _fp(U), _fp(V))
        return U, V

    def initialize(self, ds):
        lib().orc_model_initialize(self.h, ds.h)

    # hooks for the CPU emulation of the row-sharded epoch
    def set_range(self, lo, hi):
        lib().orc_model_set_range(self.h, int(lo), int(hi))

    def put_factors(self, U=None, V=None):
        U = None if U is None else np.ascontiguousarray(U, np.float32)
        V = None if V is None else np.ascontiguousarray(V, np.float32)
        lib().orc_model_put_factors(self.h, _fp(U), _fp(V))

    def set_item_gramian(self, G):
        G = np.ascontiguousarray(G, np.float32)
        lib().orc_model_set_item_gramian(self.h, _fp(G))

    def set_gz_override(self, G):
        G = None if G is None else np.ascontiguousarray(G, np.float32)
        lib().orc_model_set_gz_override(self.h, _fp(G))

    def block_step(self, ds, item_side, bs, be, pred):
        """One ++ block sweep on the rows in range; `pred` (float32 [num_tuples]) is updated in place."""
        assert pred.dtype == np.float32 and pred.flags.c_contiguous
        lib().orc_model_block_step(self.h, ds.h, int(item_side), int(bs), int(be), _fp(pred))

    def predict_rows(self, ds, item_side, pred):
        """pred[t] = x_row . e_col for the tuples of the rows in range."""
        assert pred.dtype == np.float32 and pred.flags.c_contiguous
        lib().orc_model_predict_rows(self.h, ds.h, int(item_side), _fp(pred))

    def train(self, ds):
        lib().orc_model_train(self.h, ds.h)

    def stage(self, ds, stage):
        lib().orc_model_stage(self.h, ds.h, stage)

    def set_print_train_stats(self, on):
        lib().orc_model_set_print_train_stats(self.h, int(on))

    def state(self):
        z = np.zeros(self.num_users, np.float32)
        loss = np.zeros(self.num_users, np.float32)
        hs = np.zeros(self.num_users, np.float32)
        ireg = np.zeros(self.num_items, np.float32)
        sc = np.zeros(3, np.float32)
        G = np.zeros((self.dim, self.dim), np.float32)
        lib().orc_model_get_state(self.h, _fp(z), _fp(loss), _fp(hs), _fp(ireg), _fp(sc), _fp(G))
        return dict(z=z, loss=loss, hist_size=hs, item_reg=ireg, xi=float(sc[0]),
                    weighted_loss=float(sc[1]), mean_weight=float(sc[2]), gramian=G)

    def set_state(self, z=None, loss=None, xi=0.0):
        z = None if z is None else np.ascontiguousarray(z, np.float32)
        loss = None if loss is None else np.ascontiguousarray(loss, np.float32)
        lib().orc_model_set_state(self.h, _fp(z), _fp(loss), float(xi))

    def stats(self, ds=None):
        out = (C.c_double * 6)()
        if ds is None:
            lib().orc_model_get_stats(self.h, out)
        else:
            lib().orc_model_compute_stats(self.h, ds.h, out)
        keys = ["loss", "loss_observed", "loss_unobserved", "loss_reg", "loss_reg_user", "loss_reg_item"]
        return dict(zip(keys, list(out)))

    def last_snr(self):
        ni, ns = C.c_int(), C.c_int()
        lib().orc_model_last_snr(self.h, C.byref(ni), C.byref(ns), None)
        out = np.zeros((ni.value, ns.value), np.int32)
        if out.size:
            lib().orc_model_last_snr(self.h, C.byref(ni), C.byref(ns), _ip(out))
        return out

    def evaluate(self, tr, te, k_list=(5, 10, 20, 50, 100), want_topk=False, want_folded=False):
        ks = np.asarray(k_list, np.int32)
        nu = lib().orc_model_evaluate(self.h, tr.h, te.h, _ip(ks), len(ks), None, None, None, None, None)
        ids = np.zeros(nu, np.int32)
        rec = np.zeros((nu, len(ks)), np.float32)
        ndcg = np.zeros((nu, len(ks)), np.float32)
        topk = np.zeros((nu, int(ks.max())), np.int32) if want_topk else None
        folded = np.zeros((nu, self.dim), np.float32) if want_folded else None
        lib().orc_model_evaluate(self.h, tr.h, te.h, _ip(ks), len(ks), _ip(ids), _fp(rec), _fp(ndcg),
                                 _ip(topk), _fp(folded))
        return dict(user_ids=ids, recall=rec, ndcg=ndcg, topk=topk, folded=folded)


def gramian(E, w=None):
    E = np.ascontiguousarray(E, np.float32)
    w = None if w is None else np.ascontiguousarray(w, np.float32)
    out = np.zeros((E.shape[1], E.shape[1]), np.float32)
    lib().orc_gramian(_fp(E), E.shape[0], E.shape[1], _fp(w), _fp(out))
    return out


def init_factors(nu, ni, d, stdev=0.1, seed=12345):
    U = np.zeros((nu, d), np.float32)
    V = np.zeros((ni, d), np.float32)
    lib().orc_init_factors(nu, ni, d, stdev, seed, _fp(U), _fp(V))
    return U, V


def metric_cvar(ms, alphas):
    ms = np.ascontiguousarray(ms, np.float32)
    alphas = np.ascontiguousarray(alphas, np.float32)
    out = np.zeros(len(alphas), np.float32)
    lib().orc_metric_cvar(_fp(ms), len(ms), _fp(alphas), len(alphas), _fp(out))
    return out


def num_threads():
    return lib().orc_num_threads()
