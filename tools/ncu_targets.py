"""Launch ONE instance of a kernel of interest at its benchmark shape, for `ncu --set full` captures (B200_PROFILING.md).
usage: ncu --set full --import-source on --clock-control none -k regex:<kernel> -c <n> -o gpurun_out/<name> \
           python tools/ncu_targets.py <target>
targets: sim (exact similarity, 8 x 10240 x 20480 x 64), sim128 (8 x 1280 x 1280 x 128, the in-step shape),
         gemm (tf32 163840 x 128 x 32, the dominant HBM-bound shape), gemmx3 (3xTF32 163840 x 128 x 64 and 10240 x 1024 x 2048),
         knn (pyramid tables of 8 frames)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from cofii2p_b200 import ops

t = sys.argv[1]
torch.manual_seed(0)
if t in ("sim", "sim128"):
    npt, npx, c, frames = (10240, 20480, 64, 8) if t == "sim" else (1280, 1280, 128, 8)
    pt, pth = ops.l2norm_rows_f16(torch.randn(frames * npt, c, device="cuda"))
    px, pxh = ops.l2norm_rows_f16(torch.randn(frames * npx, c, device="cuda"))
    for _ in range(2):
        ops.sim_argmin(pt, px, frames, pt_h=pth, px_h=pxh)
        ops.sim_argmin_f16(pth, pxh, frames)
elif t == "gemm":
    ops.set_engine("tf32")
    a, w, b = torch.randn(163840, 32, device="cuda"), torch.randn(128, 32, device="cuda"), torch.randn(128, device="cuda")
    for _ in range(2):
        ops.gemm(a, w, bias=b)
    a2, w2 = torch.randn(20480, 128, device="cuda"), torch.randn(128, 128, device="cuda")
    for _ in range(2):
        ops.gemm(a2, w2)
elif t == "gemmx3":
    ops.set_engine("tf32x3")
    for (m, n, k) in ((163840, 128, 64), (10240, 1024, 2048), (20480, 128, 128)):
        a, w = torch.randn(m, k, device="cuda"), torch.randn(n, k, device="cuda")
        for _ in range(2):
            ops.gemm(a, w)
elif t == "gemmx3s":   # persistent 3xTF32 kernel with pre-split weights (csrc/gemm_x3.cu): HBM-bound and tensor-bound shapes
    ops.set_engine("tf32x3")
    for (m, n, k) in ((163840, 128, 32), (20480, 1024, 3072), (20480, 128, 128)):
        a, w, b = torch.randn(m, k, device="cuda"), torch.randn(n, k, device="cuda"), torch.randn(n, device="cuda")
        for _ in range(2):
            ops.gemm(a, w, bias=b, const_w=True)
elif t == "knn":
    from cofii2p_b200.frames import make_frame, stack_frames
    batch = stack_frames([make_frame(i, cache_dir="/tmp/cofi_frames", device="cuda") for i in range(8)])
    pts = [p.cuda() for p in batch["pc_data_dict"]["points"]]
    for _ in range(2):
        ops.knn_pyramid(pts, frames=8, k=128, k_up=1)
torch.cuda.synchronize()
ops.set_engine("fp32")
