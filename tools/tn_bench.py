"""Micro-benchmark of the weight-gradient contraction (cofi_gemm_tn / cofi_conv2d_wgrad_nhwc) on both engines."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cofii2p_b200 import ops  # noqa: E402


def timeit(fn, n=10):
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
    shapes = [(81920, 960, 64), (81920, 64, 64), (40960, 1920, 128), (20480, 3840, 256), (10240, 7680, 512), (5120, 15360, 1024),
              (81920, 64, 256), (20480, 256, 512), (5120, 128, 128)]
    for R, Mo, No in shapes:
        a, b = torch.randn((R, Mo), device="cuda"), torch.randn((R, No), device="cuda")
        row = []
        for eng in ("fp32", "tf32"):
            ops.set_engine(eng)
            ms = timeit(lambda: ops.gemm_tn(a, b))
            row.append(f"{eng} {ms * 1e3:8.1f} us {2.0 * R * Mo * No / ms / 1e9:7.1f} TF/s {4.0 * (R * Mo + R * No) / ms / 1e6:7.0f} GB/s")
        print(f"gemm_tn R={R:6d} Mo={Mo:5d} No={No:5d} | " + " | ".join(row), flush=True)
    for (B, H, W, Cin, Cout, k, s, p) in [(4, 80, 256, 64, 64, 3, 1, 1), (4, 40, 128, 64, 64, 3, 1, 1), (4, 40, 128, 64, 128, 3, 2, 1),
                                          (4, 20, 64, 128, 128, 3, 1, 1), (4, 80, 256, 128, 64, 3, 1, 1), (4, 40, 128, 256, 128, 3, 1, 1)]:
        x = torch.randn((B, H, W, Cin), device="cuda")
        Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        dy = torch.randn((B, Ho, Wo, Cout), device="cuda")
        row = []
        for eng in ("fp32", "tf32"):
            ops.set_engine(eng)
            ms = timeit(lambda: ops.conv2d_wgrad_nhwc(x, dy, k, k, s, p))
            row.append(f"{eng} {ms * 1e3:8.1f} us {2.0 * B * Ho * Wo * Cout * k * k * Cin / ms / 1e9:7.1f} TF/s")
        print(f"wgrad B={B} {H}x{W} {Cin}->{Cout} k{k} s{s} | " + " | ".join(row), flush=True)


if __name__ == "__main__":
    main()
