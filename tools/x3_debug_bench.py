import os, sys, json
sys.path.insert(0, "/root/repo")
import torch
from cofii2p_b200 import ops
ops.set_engine("tf32x3")
m, n, k = 163840, 128, 32
copies = 4
a = [torch.randn(m, k, device="cuda") for _ in range(copies)]
o = [torch.empty(m, n, device="cuda") for _ in range(copies)]
w = torch.randn(n, k, device="cuda"); b = torch.randn(n, device="cuda")
for stats in (False, True):
    f = (lambda i: ops.gemm_colstats(a[i], w, bias=b, const_w=True)) if stats else (lambda i: ops.gemm(a[i], w, bias=b, out=o[i], const_w=True))
    for i in range(copies): f(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(32): f(i % copies)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print(json.dumps({"dbg": os.environ.get("COFI_X3_DEBUG", "0"), "colstats": stats, "us": round(e0.elapsed_time(e1) / 32 * 1e3, 2)}), flush=True)
