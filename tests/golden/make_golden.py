"""Generates tests/golden/oracle_fixture_golden.json: summary numbers of the CPU
oracle on the reference's bundled ML-1M fixture under the reference's own test
settings.  The reference holds no golden vectors and cannot be built
here (Eigen absent), so these pin the ORACLE against regressions; its agreement with
the reference is pinned only by the reference's thresholds (tests/test_oracle.py)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import helpers  # noqa: E402
from oracle import loader as O  # noqa: E402
from test_oracle import CASES  # noqa: E402

tr = O.Dataset.from_csv(helpers.fixture_csv("train"))
vtr = O.Dataset.from_csv(helpers.fixture_csv("validation_tr"))
vte = O.Dataset.from_csv(helpers.fixture_csv("validation_te"))
out = {}
for case, (epochs, cfg) in sorted(CASES.items()):
    name = case if case in O.MODEL_IDS else case.rsplit("_", 1)[0]
    m = O.Model(tr.max_user + 1, tr.max_item + 1, init_seed=1, model=name, dim=8, **cfg)
    m.initialize(tr)
    for _ in range(epochs):
        m.train(tr)
    ev = m.evaluate(vtr, vte)
    s = m.state()
    out[case] = dict(ndcg20=float(ev["ndcg"][:, 2].mean()), recall20=float(ev["recall"][:, 2].mean()),
                     ndcg100=float(ev["ndcg"][:, 4].mean()), xi=s["xi"], mean_weight=s["mean_weight"],
                     weighted_loss=s["weighted_loss"])
    print(case, out[case])
m = O.Model(tr.max_user + 1, tr.max_item + 1, init_seed=1, model="safer2", dim=8, use_snr=1,
            sampling_ratio=0.5, snr_seed=123, bandwidth=0.15, uobs_weight=0.004, reg=0.004)
m.initialize(tr)
out["snr_first8_seed123"] = m.last_snr()[0, :8].tolist()
json.dump(out, open(os.path.join(HERE, "oracle_fixture_golden.json"), "w"), indent=1, sort_keys=True)
