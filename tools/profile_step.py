"""One eager batched step inside a cudaProfilerStart/Stop range (for ncu --profile-from-start off)."""
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
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
ops.set_engine(a.engine)
dev = torch.device("cuda", 0)
model, _ = build_model(dev)
batch = stack_frames([make_frame(i, cache_dir="/tmp/cofi_frames", device="cuda") for i in range(a.batch)])
eng = InferenceEngine(model, batch, use_graph=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
with torch.no_grad(), torch.cuda.stream(eng.stream):
    for _ in range(a.steps):
        eng._step_eager()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled", a.steps, "step(s), launches/step", eng.launches_per_step)
