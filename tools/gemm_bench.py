"""Micro-benchmark of cofi_gemm shapes per engine.  Each shape is launched `reps` times back to back over rotating operand /
output buffers (one CUDA graph replay) whose total footprint exceeds the 126 MB L2 (so every launch streams from HBM and no host launch latency
sits between the events); CUDA events on the launching stream.   usage: python tools/gemm_bench.py [engine ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cofii2p_b200 import ops

SHAPES = [  # (M, N, K, where it occurs in an 8-frame step)
    (20480, 128, 128, "transformer q/k/merge over both streams"), (20480, 256, 128, "transformer mlp[0] half"),
    (20480, 128, 256, "transformer mlp[2]"), (10240, 128, 128, "transformer per-stream k / merge"),
    (163840, 32, 64, "e1_2.unary1"), (163840, 128, 32, "e1_2.unary2"), (163840, 128, 64, "e1_2.shortcut"),
    (81920, 64, 128, "e2_2.unary1"), (81920, 256, 64, "e2_2.unary2"), (40960, 512, 128, "e3_2.unary2"),
    (20480, 1024, 256, "e4_2.unary2"), (10240, 2048, 512, "e5_2.unary2"), (10240, 1024, 2048, "pc_feature_layer[0]"),
    (163840, 32, 480, "e1_2 KPConv weight-apply"), (81920, 64, 960, "e2_2 KPConv weight-apply"),
    (20480, 256, 3840, "e4_2 KPConv weight-apply"), (10240, 512, 7680, "e5_2 KPConv weight-apply"),
]
engines = sys.argv[1:] or ["tf32", "tf32x3"]
for eng in engines:
    # "tf32x3s" = the persistent 3xTF32 kernel with pre-split weights (csrc/gemm_x3.cu; what the model's Linear layers run
    # under the parity preset); "tf32x3" = the one-tile-per-CTA kernel that splits both operands in shared memory
    const_w = eng == "tf32x3s"
    ops.set_engine("tf32x3" if const_w else eng)
    for (m, n, k, where) in SHAPES:
        per = 4.0 * (m * k + m * n)
        copies = max(2, int(400e6 // per) + 1)
        a = [torch.randn(m, k, device="cuda") for _ in range(copies)]
        o = [torch.empty(m, n, device="cuda") for _ in range(copies)]
        w = torch.randn(n, k, device="cuda")
        b = torch.randn(n, device="cuda")
        for i in range(copies):
            ops.gemm(a[i], w, bias=b, out=o[i], const_w=const_w)
        reps = max(copies, 30)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()          # replayed as a graph: no host launch latency between the kernels
        with torch.cuda.graph(g):
            for i in range(reps):
                ops.gemm(a[i % copies], w, bias=b, out=o[i % copies], const_w=const_w)
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / reps
        del g
        gb = 4.0 * (m * k + n * k + m * n) / 1e9
        print(json.dumps(dict(engine=eng, M=m, N=n, K=k, where=where, us=round(t * 1e3, 2), tflops=round(2.0 * m * n * k / t / 1e9, 1),
                              gbs=round(gb / t * 1e3, 1), dbg=os.environ.get("COFI_TC_DEBUG", "0"))), flush=True)
        del a, o
        torch.cuda.empty_cache()
ops.set_engine("fp32")
