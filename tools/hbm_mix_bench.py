"""Achievable HBM bandwidth as a function of the read : write mix (torch kernels over 1 GiB buffers, CUDA events, best of 5).
The small-K contractions of the point branch write 4x what they read (163840 x 128 x 32: 21 MB in, 84 MB out); measured on B200:
a pure write stream reaches 6.8 TB/s, i.e. the write-heavy mix is NOT what holds those kernels at 3.4 TB/s."""
import json
import torch

n = 256 * 1024 * 1024   # fp32 elements = 1 GiB
a = torch.empty(n, device="cuda", dtype=torch.float32).normal_()
b = torch.empty(n, device="cuda", dtype=torch.float32)


def best(fn, nbytes, reps=5):
    t = []
    for _ in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return nbytes / min(t[2:]) / 1e6


res = {
    "write_only_fill_GBps": best(lambda: b.fill_(1.0), 4 * n),
    "read_only_sum_GBps": best(lambda: a.sum(), 4 * n),
    "copy_1r_1w_GBps": best(lambda: b.copy_(a), 8 * n),
    "add_2r_1w_GBps": best(lambda: torch.add(a, b, out=b), 12 * n),
}
print(json.dumps(res))
