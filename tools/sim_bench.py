"""Fused similarity + arg-min throughput (BASELINE sweep point 10240 x 20480 x 64 and the coarse 1280 x 1280 x 128 case)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from cofii2p_b200 import ops
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
out = []
for (npt, npx, c, frames) in [(10240, 20480, 64, 1), (10240, 20480, 64, 8), (1280, 1280, 128, 8), (20480, 20480, 128, 4)]:
    pt = F.normalize(torch.randn(frames * npt, c, device="cuda"), dim=1)
    px = F.normalize(torch.randn(frames * npx, c, device="cuda"), dim=1)
    pth, pxh = ops.cast_f16(pt), ops.cast_f16(px)
    for eng, name in ((None, "tcgen05-f16"), (ops.ENGINE_TF32, "tcgen05-tf32"), (ops.ENGINE_FP32, "simt-fp32")):
        if name == "simt-fp32" and frames * npt * npx > 4e9:
            continue
        run = (lambda: ops.sim_argmin_f16(pth, pxh, frames)) if eng is None else (lambda: ops.sim_argmin(pt, px, frames, engine=eng))
        for _ in range(2): run()
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        fl = 2.0 * frames * npt * npx * c
        rec = dict(engine=name, npt=npt, npx=npx, c=c, frames=frames, us=t * 1e3, tflops=fl / t / 1e9,
                   matrix_bytes_never_written=4.0 * frames * npt * npx)
        out.append(rec); print(json.dumps(rec))
