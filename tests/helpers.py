"""Shared test helpers: package loader, fixture paths, synthetic interactions."""
import gzip
import importlib.util
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
PKG_DIR = os.path.join(ROOT, "safer2-recommender_b200")


def load_pkg():
    """Import the product package (its directory name has a hyphen)."""
    name = "safer2_recommender_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def fixture_csv(name):
    """Path of tests/golden/ml-1m/<name>.csv (the reference's bundled ML-1M fixture,
    tests/ml-1m/*.csv, stored gzipped); decompressed on first use."""
    dst = os.path.join(GOLDEN, "ml-1m", name + ".csv")
    if not os.path.exists(dst):
        with gzip.open(dst + ".gz", "rb") as f, open(dst + ".tmp", "wb") as g:
            shutil.copyfileobj(f, g)
        os.replace(dst + ".tmp", dst)
    return dst


def synth_tuples(num_users, num_items, mean_hist, seed, heavy_rows=(), empty_users=(), empty_items=()):
    """Seeded synthetic interactions: log-normal history sizes, Zipf-like item
    popularity, no duplicate (user,item) pairs, grouped by user in a random user
    order with items unsorted inside a user (like the fixture).  heavy_rows =
    [(user, n)] forces specific history lengths (n=128,129,255,...)."""
    rng = np.random.default_rng(seed)
    sizes = np.clip(np.round(np.exp(rng.normal(np.log(mean_hist), 0.9, num_users))), 1, num_items // 2).astype(int)
    for u, n in heavy_rows:
        sizes[u] = min(n, num_items)
    for u in empty_users:
        sizes[u] = 0
    allowed = np.setdiff1d(np.arange(num_items), np.asarray(empty_items, dtype=int))
    p = 1.0 / (np.arange(len(allowed)) + 5.0) ** 0.8
    p /= p.sum()
    users, items = [], []
    for u in rng.permutation(num_users):
        n = min(sizes[u], len(allowed))
        if n == 0:
            continue
        it = allowed[rng.choice(len(allowed), size=n, replace=False, p=p)]
        users.append(np.full(n, u, np.int32))
        items.append(it.astype(np.int32))
    return np.concatenate(users), np.concatenate(items)


def rel_fro(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
