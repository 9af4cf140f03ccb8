"""Per-call split of one eager batched inference step (CUDA events around every C-ABI call, algorithmic flops / bytes as
annotated in ops.py): the slowest calls with their achieved TFLOP/s and GB/s."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import build_model  # noqa: E402
from cofii2p_b200 import ops  # noqa: E402
from cofii2p_b200.engine import InferenceEngine  # noqa: E402
from cofii2p_b200.frames import make_frame, stack_frames  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--engine", default="tf32")
ap.add_argument("--top", type=int, default=40)
a = ap.parse_args()
ops.set_engine(a.engine)
dev = torch.device("cuda", 0)
model, _ = build_model(dev)
model.fork_image_stream = False
batch = stack_frames([make_frame(i, cache_dir="/tmp/cofi_frames", device="cuda") for i in range(a.batch)])
eng = InferenceEngine(model, batch, use_graph=False)
with torch.no_grad(), torch.cuda.stream(eng.stream):
    eng._step_eager()
    torch.cuda.synchronize()
    ops.profile_start()
    eng._step_eager()
    rec = ops._prof
    ops._prof = None
    torch.cuda.synchronize()
rows = [(e0.elapsed_time(e1) * 1e3, i, name, fl, by) for i, (name, e0, e1, fl, by) in enumerate(rec)]
tot = sum(r[0] for r in rows)
print(f"{len(rows)} calls, {tot / 1e3:.3f} ms")
for us, i, name, fl, by in sorted(rows, reverse=True)[:a.top]:
    print(f"#{i:4d} {name:28s} {us:8.1f} us  {fl / 1e9:9.2f} GFLOP {by / 1e6:8.1f} MB  {fl / us / 1e6:8.1f} TFLOP/s {by / us / 1e3:8.1f} GB/s")
