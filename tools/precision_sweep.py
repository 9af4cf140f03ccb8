#!/usr/bin/env python
"""Which stages of the forward need more than tf32?  Per-group precision sweep against the REAL reference's golden outputs
(tests/golden/frame_s0_n20480.npz, frozen by oracle/make_golden.py), plus the speed of every configuration.

The model tags its stages with ops.group(...): img_encoder, img_decoder, kpconv (aggregate + weight-apply), pc_unary,
pc_maxpool, pc_feature, transformer (projections / MLP), tr_attn (attention kernel), score.  For every group g the sweep runs
  "tf32 + g@x3": everything tf32, g in 3xTF32 (fp32-grade; for kpconv / pc_maxpool / tr_attn this also drops the fp16 operand
                 shortcut of that stage)        -> how much of the error g is responsible for
  "x3 + g@tf32": everything 3xTF32, g in tf32  -> the error g alone introduces
and prints one JSON line per configuration: max relative error of the six outputs, correspondences identical or not,
device-resident frames/s of the 8-frame test-mode graph.   usage: python tools/precision_sweep.py [--fast] > out.jsonl"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from cofii2p_b200 import ops  # noqa: E402
from cofii2p_b200.engine import InferenceEngine  # noqa: E402
from cofii2p_b200.frames import make_frame, stack_frames  # noqa: E402

GROUPS = ["img_encoder", "img_decoder", "kpconv", "pc_unary", "pc_maxpool", "pc_feature", "transformer", "tr_attn", "score"]


def main():
    fast = "--fast" in sys.argv
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ops.set_engine("tf32")
    model, _ = bench.build_model(dev)
    batch = stack_frames([make_frame(i, cache_dir="/tmp/cofi_frames", device="cuda") for i in range(8)])
    configs = [("tf32", "tf32", {}), ("tf32x3", "tf32x3", {})]
    if "--only-policies" not in sys.argv:
        configs.append(("fp32", "fp32", {}))
        for g in GROUPS:
            configs.append((f"tf32 + {g}@x3", "tf32", {g: "tf32x3"}))
        for g in GROUPS:
            configs.append((f"x3 + {g}@tf32", "tf32x3", {g: "tf32"}))
    for extra in sys.argv[1:]:
        if extra.startswith("--policy="):  # e.g. --policy=tf32:kpconv=tf32x3,score=tf32x3
            base, _, rest = extra[len("--policy="):].partition(":")
            pol = dict(kv.split("=") for kv in rest.split(",") if kv)
            configs.append((extra[len("--policy="):], base, pol))
    for name, base, pol in configs:
        ops.set_engine(base)
        ops.set_policy(pol)
        rec = {"config": name, "engine": base, "policy": pol, **bench.golden_parity(model, dev)}
        if not fast:
            eng = InferenceEngine(model, batch, mode="test", use_graph=True)
            for _ in range(3):
                eng.run()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            with torch.cuda.stream(eng.stream):
                e0.record(eng.stream)
                for _ in range(10):
                    eng.run()
                e1.record(eng.stream)
            torch.cuda.synchronize()
            rec["ms_per_step"] = e0.elapsed_time(e1) / 10
            rec["frames_per_s"] = 8 * 10 / (e0.elapsed_time(e1) / 1e3)
            rec["launches_per_step"] = eng.launches_per_step
            del eng
            torch.cuda.empty_cache()
        print(json.dumps(rec), flush=True)
    ops.set_engine("fp32")


if __name__ == "__main__":
    main()
