"""CPU test of the reporting side of the evaluation in the C++ host mirror (include/frecsys/evaluation.h,
SURVEY.md 8f-3): EvaluationResult::cvar() and the log lines of show() against a numpy restatement of the
reference's loop (evaluation.h:61-102: ascending sort, running float sum, position int(size * alpha) evaluated in
float, the `counter` quirk that writes results in encounter order)."""
import os
import re
import subprocess

import numpy as np

import helpers

SRC = r'''
#include <cstdio>
#include <vector>
#include "frecsys/evaluation.h"
int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb");
  int hdr[3];  // users, nk, na
  if (fread(hdr, sizeof(int), 3, f) != 3) return 1;
  const int nu = hdr[0], nk = hdr[1], na = hdr[2];
  frecsys::VectorXi k(nk); frecsys::VectorXf a(na);
  if (fread(k.data(), sizeof(int), nk, f) != (size_t)nk) return 1;
  if (fread(a.data(), sizeof(float), na, f) != (size_t)na) return 1;
  frecsys::MatrixXf rec(nu, nk), ndcg(nu, nk);
  if (fread(rec.data(), sizeof(float), (size_t)nu * nk, f) != (size_t)nu * nk) return 1;
  if (fread(ndcg.data(), sizeof(float), (size_t)nu * nk, f) != (size_t)nu * nk) return 1;
  fclose(f);
  frecsys::EvaluationResult r{k, a, rec, ndcg};
  for (int i = 0; i < nk; ++i) {
    frecsys::VectorXf c = r.cvar(r.recall.col(i));
    for (int j = 0; j < na; ++j) printf("CVAR %d %d %.9g\n", i, j, c[j]);
  }
  r.show();
  return 0;
}
'''


def _ref_cvar(col, alphas):
    ms = np.sort(col.astype(np.float32))
    out = np.zeros(len(alphas), np.float32)
    counter = 0
    acc = np.float32(0)
    for i, v in enumerate(ms):
        acc = np.float32(acc + v)
        for j in range(counter, len(alphas)):
            pos = int(np.float32(len(ms)) * np.float32(alphas[j]))   # size_t * float -> float, truncated
            if pos == i:
                out[counter] = np.float32(acc / np.float32(i + 1))
                counter += 1
    return out


def test_cvar_and_log_lines(tmp_path):
    exe = os.path.join(tmp_path, "evalrep")
    src = os.path.join(tmp_path, "evalrep.cc")
    open(src, "w").write(SRC)
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(helpers.ROOT, "include"), "-o", exe, src], check=True)
    rng = np.random.default_rng(9)
    nu = 997
    ks = np.array([5, 20, 100], np.int32)
    alphas = np.arange(1, 10, dtype=np.float32) / np.float32(10)   # run_model.cc:226-232: 0.1 .. 0.9
    rec = rng.random((nu, 3)).astype(np.float32) * (rng.random((nu, 3)) > 0.3)
    ndcg = rng.random((nu, 3)).astype(np.float32)
    path = os.path.join(tmp_path, "in.bin")
    with open(path, "wb") as f:
        f.write(np.array([nu, 3, len(alphas)], np.int32).tobytes())
        f.write(ks.tobytes()); f.write(alphas.tobytes()); f.write(rec.tobytes()); f.write(ndcg.tobytes())
    p = subprocess.run([exe, path], check=True, capture_output=True, text=True)
    got = np.zeros((3, len(alphas)), np.float32)
    for line in p.stdout.splitlines():
        m = re.match(r"CVAR (\d+) (\d+) (\S+)", line)
        if m:
            got[int(m.group(1)), int(m.group(2))] = np.float32(m.group(3))
    for i in range(3):
        want = _ref_cvar(rec[:, i], alphas)
        np.testing.assert_allclose(got[i], want, rtol=2e-6, atol=1e-7)
        assert np.all(np.diff(got[i]) >= -1e-6)     # lower-tail means grow with the quantile
    log = p.stderr
    # the reference's lines (evaluation.h:61-81), 4 decimals
    mean_rec = " ".join(f"Mean Rec@{k}={rec[:, i].astype(np.float32).mean(dtype=np.float32):.4f}" for i, k in enumerate(ks))
    assert re.search(r"Mean Rec@5=\d\.\d{4} Mean Rec@20=\d\.\d{4} Mean Rec@100=\d\.\d{4}", log), log[:400]
    assert re.search(r"Mean NDCG@5=\d\.\d{4} Mean NDCG@20=", log)
    for q in alphas:
        assert f"Rec CVaR (q={q:.2f})@5=" in log and f"NDCG CVaR (q={q:.2f})@5=" in log
    # values in the log are the cvar() values
    w = _ref_cvar(rec[:, 0], alphas)
    assert f"Rec CVaR (q=0.10)@5={w[0]:.4f}" in log
    assert f"Rec CVaR (q=0.90)@5={w[8]:.4f}" in log
    # the means agree to the printed precision with a float mean (summation order may differ in the last digit)
    m0 = re.search(r"Mean Rec@5=(\d\.\d{4})", log)
    assert abs(float(m0.group(1)) - float(rec[:, 0].mean())) < 2e-4, (m0.group(1), mean_rec)
