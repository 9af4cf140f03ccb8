"""Where the persistent 3xTF32 kernel (csrc/gemm_x3.cu) waits: per-role barrier wait clocks per k-block for a few shapes.
usage: COFI_X3_PROFILE=1 python tools/x3_profile.py"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("COFI_X3_PROFILE", "1")
import torch

from cofii2p_b200 import lib, ops

ops.set_engine("tf32x3")
L = lib.load()
names = ["tma.a_empty", "tma.b_empty", "-", "mma.b_full", "mma.ta_full", "mma.acc_empty", "epi.acc_full", "-", "-",
         "split.a_full", "split.ta_empty", "-", "-", "ctas", "kblocks", "cta_clocks"]
shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]] or [(10240, 512, 7680), (20480, 1024, 3072), (10240, 1024, 2048),
                                                                           (163840, 128, 32), (20480, 128, 128)]
for (m, n, k) in shapes:
    a, w = torch.randn(m, k, device="cuda"), torch.randn(n, k, device="cuda")
    ops.gemm(a, w, const_w=True)
    buf = (ctypes.c_ulonglong * 16)()
    L.cofi_debug_x3_profile(buf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.gemm(a, w, const_w=True)
    e1.record()
    torch.cuda.synchronize()
    L.cofi_debug_x3_profile(buf)
    v = list(buf)
    ctas, kbs, clk = max(v[13], 1), max(v[14], 1), v[15]
    rec = {"shape": f"{m}x{n}x{k}", "us": round(e0.elapsed_time(e1) * 1e3, 1), "ctas": ctas, "kblocks_per_cta": kbs / ctas,
           "clk_per_kblock": round(clk / kbs, 1)}
    for i, nm in enumerate(names[:11]):
        if nm != "-":
            rec[nm + "_clk_per_kb"] = round(v[i] / kbs, 1)
    print(json.dumps(rec), flush=True)
