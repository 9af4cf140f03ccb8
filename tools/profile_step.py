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
ap.add_argument("--mode", default="test")
ap.add_argument("--emit-calls", default=None, help="write [(op family key as bench.py names it, kernels launched)] of the profiled step")
a = ap.parse_args()
ops.set_engine(a.engine)
dev = torch.device("cuda", 0)
model, _ = build_model(dev)
batch = stack_frames([make_frame(i, cache_dir="/tmp/cofi_frames", device="cuda") for i in range(a.batch)])
eng = InferenceEngine(model, batch, mode=a.mode, use_graph=False)
model.fork_image_stream = False   # one stream: the launch order of the ncu list is the call order
with torch.no_grad(), torch.cuda.stream(eng.stream):
    eng._step_eager()
torch.cuda.synchronize()
if a.emit_calls:
    import json
    from bench import measured_peaks
    hbm, _, tf_sust, _ = measured_peaks()
    ops.profile_start()
torch.cuda.cudart().cudaProfilerStart()
with torch.no_grad(), torch.cuda.stream(eng.stream):
    for _ in range(a.steps):
        eng._step_eager()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
if a.emit_calls:
    _, calls = ops.profile_stop(ridge=0.5 * tf_sust * 1e12 / (hbm * 1e9), raw=True)
    json.dump(calls, open(a.emit_calls, "w"))
print("profiled", a.steps, "step(s), launches/step", eng.launches_per_step)
