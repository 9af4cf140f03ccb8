#!/usr/bin/env python
"""Pyramid/KNN-128 table builder timing (csrc/knn.cu): all 13 tables of B stacked synthetic
KITTI frames, CUDA events, plus the brute-force (no culling) variant and the torch dense-matrix + topk the reference's
`precompute_point_cloud_cuda` runs (model/kpconv/preprocess_data.py:131-143) for one frame."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cofii2p_b200 import ops  # noqa: E402
from cofii2p_b200.frames import make_frame  # noqa: E402


def timed(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--num-pc", type=int, default=20480)
    args = ap.parse_args()
    fr = [make_frame(s, num_pc=args.num_pc, cache_dir="/tmp/cofi_frames", device="cuda") for s in range(min(args.frames, 2))]
    B = args.frames
    levels = [torch.cat([fr[b % len(fr)]["pc_data_dict"]["points"][l] for b in range(B)], 0).cuda() for l in range(5)]
    n = [l.shape[0] // B for l in levels]
    pairs = sum(x * x for x in n) + 2 * sum(n[i] * n[i + 1] for i in range(4))
    out_bytes = 8 * 128 * (sum(n) + sum(n[1:]) + sum(n[:-1]))
    ws = ops._ws(ops._lib.cofi_knn_pyramid_workspace((__import__("ctypes").c_int64 * 5)(*n), 5, B), levels[0].device)
    res = {"frames": B, "num_pc": args.num_pc, "pairs_per_frame": pairs, "table_bytes_per_frame": out_bytes}
    for name, mode in (("direct", 0), ("expanded", 1), ("direct_iterative_selection", ops.KNN_NOFAST),
                       ("direct_engine_tables_k_up_1", -1), ("direct_nocull", ops.KNN_NOCULL)):
        if mode == -1:   # what the engine builds: up-sampling tables reduced to their single live column
            ms = timed(lambda: ops.knn_pyramid(levels, frames=B, k=128, k_up=1, mode=0, workspace=ws), 10)
            res[name] = {"ms_per_batch": ms, "ms_per_frame": ms / B}
            continue
        ms = timed(lambda: ops.knn_pyramid(levels, frames=B, k=128, mode=mode, workspace=ws), 10 if mode < 256 else 3)
        res[name] = {"ms_per_batch": ms, "ms_per_frame": ms / B, "gpairs_per_s": pairs * B / ms / 1e6,
                     "table_write_GBps": out_bytes * B / ms / 1e6}
    one = [l[: n[i]] for i, l in enumerate(levels)]
    ms1 = timed(lambda: ops.knn_pyramid(one, frames=1, k=128, mode=0), 10)
    res["direct_single_frame_ms"] = ms1

    def torch_ref():
        for s, q in [(l, l) for l in range(5)] + [(l, l + 1) for l in range(4)] + [(l + 1, l) for l in range(4)]:
            a, b = one[q], one[s]
            d = -2.0 * a @ b.t() + (a * a).sum(1, keepdim=True) + (b * b).sum(1)[None]
            d.clamp_(min=1e-12).topk(128, dim=1, largest=False)
    res["torch_dense_topk_single_frame_ms"] = timed(torch_ref, 3)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
