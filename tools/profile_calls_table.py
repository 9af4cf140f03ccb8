"""Per-call table of one eager batched step: C-ABI entry point, shape tag, microseconds (CUDA events), achieved GB/s and
TFLOP/s of the algorithmic work.   usage: python tools/profile_calls_table.py [--engine parity] [--mode test] > out.md"""
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
ap.add_argument("--engine", default="parity")
ap.add_argument("--mode", default="test")
ap.add_argument("--tables", default="host")
a = ap.parse_args()
ops.set_engine(a.engine)
dev = torch.device("cuda", 0)
model, _ = build_model(dev)
batch = stack_frames([make_frame(i, cache_dir="/tmp/cofi_frames", device="cuda") for i in range(a.batch)])
eng = InferenceEngine(model, batch, mode=a.mode, use_graph=False, tables=a.tables)
model.fork_image_stream = False
REP = 3
with torch.no_grad(), torch.cuda.stream(eng.stream):
    eng._step_eager()
    torch.cuda.synchronize()
    per = None
    for _ in range(REP):
        ops.profile_start()
        eng._step_eager()
        ops.profile_stop()
        calls = ops._prof_calls
        if per is None:
            per = [list(c) for c in calls]
        else:
            for p, c in zip(per, calls):
                p[1] = min(p[1], c[1])
tot = sum(p[1] for p in per)
print(f"# per-call table, engine={a.engine} mode={a.mode} tables={a.tables}: {len(per)} calls, {tot:.3f} ms (min of {REP} eager passes)\n")
print("| # | entry point | shape | launches | us | GB/s | TFLOP/s |\n|---|---|---|---|---|---|---|")
for i, (name, ms, nl, fl, by, tag) in enumerate(per):
    print(f"| {i} | {name} | {tag} | {nl} | {ms*1e3:.1f} | {by/ms/1e6 if ms else 0:.0f} | {fl/ms/1e9 if ms else 0:.1f} |")
