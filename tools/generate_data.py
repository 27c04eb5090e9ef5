#!/usr/bin/env python
"""Dataset generator CLI: writes the five `uid,sid` CSV files run_model reads
(train.csv, validation_tr.csv, validation_te.csv, test_tr.csv, test_te.csv) for a SYNTHETIC data set with the
shape statistics of the reference's data sets.  The reference's own scripts (scripts/generate_data.py) download
MovieLens / MSD, which needs a network; this tool keeps their output format and split protocol:

* interactions come from bench.synth_interactions (log-normal history sizes, Zipf-like item popularity, grouped by
  user in random user order, items unsorted within a user -- like the bundled ML-1M fixture);
* strong generalisation: the last 2 x n_heldout users are held out (validation, then test), ids follow the training
  users (generate_data.py:118-161);
* every held-out user with at least 5 items gets a random 20 % of them moved to *_te.csv, the rest stay in
  *_tr.csv (split_train_test_proportion, generate_data.py:52-89, seed 98765).

    python tools/generate_data.py --shape ml20m --out /tmp/ml20m --heldout 10000
    tools/run_model --model_name safer2 --train_data /tmp/ml20m/train.csv \\
        --test_train_data /tmp/ml20m/test_tr.csv --test_test_data /tmp/ml20m/test_te.csv ...
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (SHAPES, synth_interactions)

HELDOUT = {"ml1m": 1000, "ml20m": 10000, "msd": 50000, "tiny": 200}  # generate_data.py:167,206,223


def write_csv(path, users, items):
    with open(path, "w") as f:
        f.write("uid,sid\n")
        np.savetxt(f, np.stack([users, items], 1), fmt="%d", delimiter=",")


def split_tr_te(users, items, rng, test_prop=0.2):
    """Per user (tuples are grouped by user): users with >= 5 items give int(0.2 n) random items to te."""
    start = np.flatnonzero(np.r_[True, users[1:] != users[:-1]])
    end = np.r_[start[1:], users.shape[0]]
    te = np.zeros(users.shape[0], bool)
    for b, e in zip(start, end):
        n = e - b
        if n >= 5:
            te[b + rng.choice(n, size=int(test_prop * n), replace=False)] = True
    return (users[~te], items[~te]), (users[te], items[te])


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--shape", default="tiny", choices=sorted(bench.SHAPES))
    ap.add_argument("--users", type=int, help="training users (overrides --shape)")
    ap.add_argument("--items", type=int)
    ap.add_argument("--nnz", type=int, help="training tuples (approximate)")
    ap.add_argument("--heldout", type=int, help="held-out users per split (validation and test)")
    ap.add_argument("--seed", type=int, default=98765)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    nu, ni, nnz = bench.SHAPES[args.shape]
    nu, ni, nnz = args.users or nu, args.items or ni, args.nnz or nnz
    nh = args.heldout if args.heldout is not None else HELDOUT.get(args.shape, max(1, nu // 14))
    total_users = nu + 2 * nh
    users, items = bench.synth_interactions(total_users, ni, int(nnz * total_users / nu), seed=args.seed)
    os.makedirs(args.out, exist_ok=True)
    rng = np.random.default_rng(args.seed)
    tr = users < nu
    write_csv(os.path.join(args.out, "train.csv"), users[tr], items[tr])
    report = {"train": int(tr.sum())}
    for name, lo in (("validation", nu), ("test", nu + nh)):
        sel = (users >= lo) & (users < lo + nh)
        (utr, itr), (ute, ite) = split_tr_te(users[sel], items[sel], rng)
        write_csv(os.path.join(args.out, f"{name}_tr.csv"), utr, itr)
        write_csv(os.path.join(args.out, f"{name}_te.csv"), ute, ite)
        report[name + "_tr"], report[name + "_te"] = int(utr.shape[0]), int(ute.shape[0])
    print(f"{args.out}: {nu} training users + 2 x {nh} held-out, {ni} items; tuples {report}")


if __name__ == "__main__":
    main()
