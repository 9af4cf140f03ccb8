"""TEST INFRASTRUCTURE ONLY -- freezes outputs of the reference's own `knn()` and of its stack-mode driver
(reference model/kpconv/preprocess_data.py:131-143 and :36-107) into tests/golden/knn_ref.npz.

Runs in the build container only (needs /root/reference through oracle/ref_shim.py).  Inputs are pure functions of
seeds so the tests rebuild them anywhere:
  * `knn` case: integer-lattice cloud (every fp32 distance exact, so the only freedom left to torch.topk is the order
    of equal distances; the tests compare distance rows and the index sets below the k-th distance);
  * driver case: np.random.seed(7) -> precompute_point_cloud_stack_mode on a 2048-point lattice cloud (the reference needs >= 128 points at the coarsest level) with the shim's
    exact KNNSearch stand-in: pins the half-sampling (np.random.choice WITH replacement) and which level queries which.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_shim import load_reference_preprocess  # noqa: E402


def knn_case():
    rng = np.random.default_rng(11)
    src = rng.integers(-200, 200, (2048, 3)).astype(np.float32)
    src[:, 1] = np.round(src[:, 1] / 20.0)  # flat scene: plenty of equal distances
    qry = np.concatenate([src[rng.permutation(2048)[:192]], rng.integers(-220, 220, (64, 3)).astype(np.float32)], 0)
    return src, qry


def driver_case():
    rng = np.random.default_rng(12)
    pts = rng.integers(-60, 60, (3, 2048)).astype(np.float32)
    pts[1] = np.round(pts[1] / 10.0)
    return pts


def main():
    pp = load_reference_preprocess()
    src, qry = knn_case()
    idx = pp.knn(torch.from_numpy(src), torch.from_numpy(qry), 128).numpy()
    pts = driver_case()
    np.random.seed(7)
    d = pp.precompute_point_cloud_stack_mode(pts, None, None, lengths=2048, num_stages=5)
    out = {"knn_idx": idx.astype(np.int16)}
    for i, p in enumerate(d["points"]):
        out[f"points{i}"] = p.numpy().astype(np.float32)
    for name in ("neighbors", "subsampling", "upsampling"):
        for i, t in enumerate(d[name]):
            out[f"{name}{i}"] = t.numpy().astype(np.int16)
    out["lengths"] = np.asarray(d["lengths"], dtype=np.int64)
    path = os.path.join(ROOT, "tests", "golden", "knn_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
