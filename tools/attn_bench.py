"""BASELINE config 4: I2P transformer cross-attention on tcgen05 -- the reference size (1280 x 1280, d_model 128, 4 heads)
and the named sweep point (1280 super-pixels x 1024 super-points, d_model 256, 4 heads)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cofii2p_b200 import ops
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
for (L, S, dm, heads, frames) in [(1280, 1280, 128, 4, 1), (1280, 1280, 128, 4, 8), (1280, 1024, 256, 4, 1), (1280, 1024, 256, 4, 8)]:
    q = torch.randn(frames * L, dm, device="cuda"); k = torch.randn(frames * S, dm, device="cuda")
    v = torch.randn(frames * S, dm, device="cuda"); vt = v.t().contiguous()
    D = dm // heads
    runs = [("tcgen05-tf32", lambda: ops.attention_vt(q, k, vt, frames, heads, 1.0 / D ** 0.5))]
    if D == 32:
        runs.append(("simt-fp32", lambda: ops.attention(q, k, v, frames, heads, 1.0 / D ** 0.5, engine=ops.ENGINE_FP32)))
    for name, run in runs:
        for _ in range(2): run()
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        print(json.dumps(dict(engine=name, L=L, S=S, d_model=dm, heads=heads, frames=frames, us=t * 1e3,
                              tflops=4.0 * frames * L * S * dm / t / 1e9)))
