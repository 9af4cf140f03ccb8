"""Neighbour max-pool (strided residual shortcut, reference model/kpconv/functional.py:53-66) at the four level transitions of
an 8-frame batch: fp16-gather kernel and fp32 kernel, CUDA events."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cofii2p_b200 import ops
from cofii2p_b200.frames import make_frame, stack_frames
B = 8
batch = stack_frames([make_frame(i, cache_dir="/tmp/cofi_frames", device="cuda") for i in range(B)])
sub = [t.cuda() for t in batch["pc_data_dict"]["subsampling"]]
n = [p.shape[0] for p in batch["pc_data_dict"]["points"]]
for l, C in enumerate((64, 128, 256, 512)):
    x = torch.randn(n[l], C, device="cuda")
    xh = ops.cast_f16(x)
    for name, fn in (("f16", lambda: ops.maxpool_rows_f16(xh, sub[l], B)), ("f32", lambda: ops.maxpool_rows(x, sub[l], B))):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10):
            fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        gb = sub[l].shape[0] * 128 * C * (2 if name == "f16" else 4) / 1e9
        print(json.dumps(dict(level=l, C=C, kernel=name, us=us, gather_GB=gb, gather_TBps=gb / us * 1e3)))
