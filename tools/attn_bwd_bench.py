"""Attention backward at the training size (4 frames x 1280 x 1280, d_model 128, 4 heads): tcgen05 two-pass kernel
(csrc/attention_bwd_tc.cu) vs the SIMT fp32 backward, CUDA events, L2 flushed between runs."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cofii2p_b200 import ops
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
for frames in (1, 4, 8):
    L = S = 1280
    q, k, v, go = (torch.randn(frames * L, 128, device="cuda") for _ in range(4))
    ops.set_engine("tf32")
    out, lse = ops.attention_fwd_lse(q, k, v, frames, 4, 32 ** -0.5)
    for name in ("tf32", "fp32"):
        ops.set_engine(name)
        run = lambda: ops.attention_bwd(q, k, v, out, go, lse, frames, 4, 32 ** -0.5)
        for _ in range(2): run()
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        print(json.dumps(dict(engine={"tf32": "tcgen05-tf32 (3 transposes + dQ pass + dK/dV pass)", "fp32": "simt-fp32"}[name],
                              L=L, S=S, frames=frames, us=t * 1e3, tflops=10.0 * frames * L * S * 128 / t / 1e9)))
ops.set_engine("fp32")
