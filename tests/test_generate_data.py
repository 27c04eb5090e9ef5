"""CPU test of tools/generate_data.py: the five `uid,sid` files of the reference's data layout
(scripts/generate_data.py:118-161), readable by the oracle's Dataset like the bundled fixture."""
import os
import subprocess
import sys

import numpy as np

import helpers
from oracle import loader as O


def test_generator_cli_writes_the_reference_layout(tmp_path):
    out = str(tmp_path / "gen")
    subprocess.run([sys.executable, os.path.join(helpers.ROOT, "tools", "generate_data.py"), "--users", "600", "--items", "300",
                    "--nnz", "20000", "--heldout", "50", "--out", out], check=True)
    names = ["train", "validation_tr", "validation_te", "test_tr", "test_te"]
    data = {}
    for n in names:
        path = os.path.join(out, n + ".csv")
        assert open(path).readline().strip() == "uid,sid"
        data[n] = np.loadtxt(path, delimiter=",", skiprows=1, dtype=np.int64).reshape(-1, 2)
    assert data["train"][:, 0].max() < 600
    for split, lo in (("validation", 600), ("test", 650)):
        tr, te = data[split + "_tr"], data[split + "_te"]
        assert tr[:, 0].min() >= lo and tr[:, 0].max() < lo + 50 and te[:, 0].min() >= lo and te[:, 0].max() < lo + 50
        # 80 / 20 per user, no pair in both files
        for u in np.unique(tr[:, 0]):
            n_tr, n_te = (tr[:, 0] == u).sum(), (te[:, 0] == u).sum()
            assert n_te == int(0.2 * (n_tr + n_te)) or n_tr + n_te < 5
        both = set(map(tuple, tr)) & set(map(tuple, te))
        assert not both
    ds = O.Dataset.from_csv(os.path.join(out, "train.csv"))
    assert ds.num_tuples == data["train"].shape[0] and ds.max_user == data["train"][:, 0].max()
