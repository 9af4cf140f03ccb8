"""Micro-benchmark of cofi_gemm shapes (CUDA events, L2 flushed between iterations)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cofii2p_b200 import ops
shapes = [(163840, 128, 32), (163840, 32, 64), (163840, 64, 480), (81920, 256, 64), (20480, 1024, 3072), (10240, 512, 7680),
          (10240, 128, 128), (10240, 256, 256)]
eng = sys.argv[1] if len(sys.argv) > 1 else "tf32"
ops.set_engine(eng)
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
for (m, n, k) in shapes:
    a = torch.randn(m, k, device="cuda"); w = torch.randn(n, k, device="cuda"); b = torch.randn(n, device="cuda")
    out = torch.empty(m, n, device="cuda")
    for _ in range(3): ops.gemm(a, w, bias=b, out=out)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm(a, w, bias=b, out=out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    gb = 4.0 * (m * k + n * k + m * n) / 1e9
    print(f"{eng} dbg={os.environ.get('COFI_TC_DEBUG','0')} M={m} N={n} K={k}: {t*1e3:8.1f} us  {2.0*m*n*k/t/1e9:8.1f} TFLOP/s  {gb/t*1e3:7.1f} GB/s")
