"""Generates tests/golden/ref_golden.npz: outputs of the REFERENCE's own headers
(/root/reference/include/frecsys/*.h, unmodified) compiled against the Eigen/glog API shim in
oracle/eigen_shim (`make -C oracle _ref`), run on the reference's bundled ML-1M fixture with injected
initial factors (mt19937 seed, recommender.h:61-67 order).  Only the dense arithmetic of the shim is
ours; control flow, stage order and every quirk are the reference's.  Runs only where /root/reference
exists; the resulting file is committed and used by tests on any box.

Stored per case: every 8th row of U and V, z / loss for every 8th user, xi, mean weights per epoch,
mean Recall@k / NDCG@k, NDCG@20 CVaR — small enough to commit."""
import os
import struct
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402

CASES = {
    # name: (model, dim, epochs, flags)
    "ials_d8_e1": ("ials", 8, 1, dict(uobs_weight=0.1, reg=0.003)),
    "ials_d8_e3": ("ials", 8, 3, dict(uobs_weight=0.1, reg=0.003)),
    "ials_d32_e1": ("ials", 32, 1, dict(uobs_weight=0.2, reg=0.006)),            # README.md:63 (config 1)
    "ialspp_d8_e1": ("ialspp", 8, 1, dict(uobs_weight=0.1, reg=0.003, block_size=4)),
    "ialspp_d8_e3": ("ialspp", 8, 3, dict(uobs_weight=0.1, reg=0.003, block_size=4)),
    "erm_mf_d8_e1": ("erm_mf", 8, 1, dict(uobs_weight=0.004, reg=0.005)),
    "erm_mf_d8_e3": ("erm_mf", 8, 3, dict(uobs_weight=0.004, reg=0.005)),
    "cvar_mf_d8_e1": ("cvar_mf", 8, 1, dict(uobs_weight=0.008, reg=0.002, stepsize=0.4)),
    "cvar_mf_d8_e3": ("cvar_mf", 8, 3, dict(uobs_weight=0.008, reg=0.002, stepsize=0.4)),
    "safer2_d8_e1": ("safer2", 8, 1, dict(uobs_weight=0.004, reg=0.004, bandwidth=0.15)),
    "safer2_d8_e3": ("safer2", 8, 3, dict(uobs_weight=0.004, reg=0.004, bandwidth=0.15)),
    "safer2_d32_e1": ("safer2", 32, 1, dict(uobs_weight=0.004, reg=0.004, bandwidth=0.15)),  # README.md:58 (config 2)
    "safer2_ep_d8_e2": ("safer2", 8, 2, dict(uobs_weight=0.004, reg=0.004, bandwidth=0.7, use_epanechnikov=1)),
    "safer2_pd2_d8_e2": ("safer2", 8, 2, dict(uobs_weight=0.004, reg=0.004, bandwidth=0.15, pd_iterations=2)),
    "safer2pp_d8_e1": ("safer2pp", 8, 1, dict(uobs_weight=0.004, reg=0.004, bandwidth=0.15, block_size=4)),
    "safer2pp_d8_e3": ("safer2pp", 8, 3, dict(uobs_weight=0.004, reg=0.004, bandwidth=0.15, block_size=4)),
    "ials_cg_d8_e1": ("ials", 8, 1, dict(uobs_weight=0.1, reg=0.003, use_cg=1)),
    # --use_cg: Eigen ConjugateGradient<Lower> (safer2.h:152-157) and ERM-MF's BiCGSTAB on the unsymmetrised
    # matrix (erm_mf.h:139-145, SURVEY B-5)
    "safer2_cg_d8_e1": ("safer2", 8, 1, dict(uobs_weight=0.004, reg=0.004, bandwidth=0.15, use_cg=1)),
    "erm_mf_cg_d8_e1": ("erm_mf", 8, 1, dict(uobs_weight=0.004, reg=0.005, use_cg=1)),
    "erm_mf_cg_d32_e1": ("erm_mf", 32, 1, dict(uobs_weight=0.004, reg=0.005, use_cg=1)),
    # d = 128: the tcgen05 row kernels (direct + dual form) and the TMA Gramian; default block_size = 64 (run_model.cc:174)
    "safer2_d128_e1": ("safer2", 128, 1, dict(uobs_weight=0.004, reg=0.004, bandwidth=0.15)),
    "ials_d128_e1": ("ials", 128, 1, dict(uobs_weight=0.2, reg=0.006)),
    "ialspp_d128_b64_e1": ("ialspp", 128, 1, dict(uobs_weight=0.1, reg=0.003, block_size=64)),
    "safer2pp_d128_b64_e1": ("safer2pp", 128, 1, dict(uobs_weight=0.004, reg=0.004, bandwidth=0.15, block_size=64)),
    # CVaR-MF at d = 128: the gradient-step variant of the tensor-core row kernel; the second epoch has z in {0, 1}
    "cvar_mf_d128_e2": ("cvar_mf", 128, 2, dict(uobs_weight=0.008, reg=0.002, stepsize=0.4)),
}


def read_blob(path):
    b = open(path, "rb").read()
    off, out = 0, []
    while off < len(b):
        n = struct.unpack_from("<Q", b, off)[0]
        off += 8
        out.append(np.frombuffer(b, np.float32, n, off).copy())
        off += 4 * n
    return out


def main():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "_ref"], check=True)
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    tr, vtr, vte = (helpers.fixture_csv(n) for n in ("train", "validation_tr", "validation_te"))
    out = {}
    for case, (model, dim, epochs, flags) in CASES.items():
        tmp = f"/tmp/ref_{case}.bin"
        args = [exe, model, tr, vtr, vte, tmp, f"dim={dim}", f"epochs={epochs}", "init_seed=1"] + [f"{k}={v}" for k, v in flags.items()]
        print(subprocess.run(args, check=True, capture_output=True, text=True).stdout.strip())
        U, V, z, loss, xi, mw, rec, ndcg, cvar = read_blob(tmp)
        U = U.reshape(-1, dim)
        V = V.reshape(-1, dim)
        out[case + "/U"] = U[::8]
        out[case + "/V"] = V[::8]
        out[case + "/z"] = z[::8]
        out[case + "/loss"] = loss[::8]
        out[case + "/xi"] = xi
        out[case + "/mean_weights"] = mw
        out[case + "/recall"] = rec
        out[case + "/ndcg"] = ndcg
        out[case + "/ndcg20_cvar"] = cvar
        out[case + "/fro"] = np.array([np.linalg.norm(U.astype(np.float64)), np.linalg.norm(V.astype(np.float64))])
    np.savez_compressed(os.path.join(HERE, "ref_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_golden.npz"), os.path.getsize(os.path.join(HERE, "ref_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
