#!/usr/bin/env python
"""bench.py -- I2P frames/s of the CoFiI2P coarse-to-fine correspondence hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA path through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: oracle port of the reference forward

A "step" is one pass of the hot path over one batch of B=8 synthetic KITTI-shaped frames per GPU
(BASELINE.json configs[1]: 3x160x512 image + 20480-point cloud with 5-level KNN-128 tables), `val`-style
forward (encoders + transformer + heads + decoder + patch/feature gathers).  Inference shards by frames: with N
GPUs every rank runs its own B frames, no data-path collective ("replicas only", weak scaling).  One JSON line
is printed by rank 0; see DESIGN.md section "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "I2P frames/sec (20480 pts, 160x512 img)"
UNIT = "frames/s"
ARGS = ("pc_data_dict", "img", "fine_center_kpt_coors", "fine_xy", "fine_pc_inline_index")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cofi", choices=["cofi", "reference"])
    ap.add_argument("--workload", default="infer", choices=["infer", "train"],
                    help="infer = BASELINE configs[1] (the headline metric); train = configs[4] (tools/train_bench.py)")
    ap.add_argument("--train-batch", type=int, default=4, help="training frames per GPU per step (configs[4])")
    ap.add_argument("--batch", type=int, default=8, help="frames per GPU per step")
    ap.add_argument("--num-pc", type=int, default=20480)
    ap.add_argument("--engine", default=os.environ.get("COFI_ENGINE", "tf32"), choices=["fp32", "tf32", "tf32x3"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-from-points", action="store_true", help="skip the leg that builds the KNN tables on the device")
    ap.add_argument("--cpu-frames", type=int, default=2, help="timed frames of the CPU baseline sample")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ helpers
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.samples, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def build_model(device):
    from cofii2p_b200.model.network import CoFiI2P
    from cofii2p_b200.options import Options_KITTI
    from cofii2p_b200.weights import seeded_state_dict
    m = CoFiI2P(Options_KITTI())
    sd = seeded_state_dict(m, 0)
    m.load_state_dict(sd, strict=True)
    return m.to(device).eval(), sd


def pick_cpu_threads(sd):
    """The reference path is many small ATen ops: on a many-core host the full core count oversubscribes badly
    (measured 68 s/frame at 128 threads vs ~4 s at 8).  Give the CPU arm its best case: probe a 4096-point frame at
    several thread counts and keep the fastest."""
    from cofii2p_b200.frames import make_frame
    from oracle import restate
    f = make_frame(200, num_pc=4096, cache_dir="/tmp/cofi_frames", device="cuda" if torch.cuda.is_available() else "cpu")
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (4, 8, 16, 32, 64, ncpu) if c <= ncpu})
    best, best_t = cands[0], float("inf")
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            restate.forward(sd, *[f[k] for k in ARGS], "val", run_dead=True)
            t = time.perf_counter()
            restate.forward(sd, *[f[k] for k in ARGS], "val", run_dead=True)
            dt = time.perf_counter() - t
            if dt < best_t:
                best, best_t = c, dt
    torch.set_num_threads(best)
    return best, cands


def cpu_forward_baseline(sd, num_pc, n_frames, mode="val"):
    """The reference forward (oracle port, incl. the dead layer3/layer4 work the reference executes) on the host
    cores at the best thread count: bounded sample of the same workload, one frame per forward as the reference runs."""
    from cofii2p_b200.frames import make_frame
    from oracle import restate
    frames = [make_frame(100 + i, num_pc=num_pc, cache_dir="/tmp/cofi_frames",
                         device="cuda" if torch.cuda.is_available() else "cpu") for i in range(n_frames + 1)]
    times = []
    with torch.no_grad():
        for i, f in enumerate(frames):
            t = time.perf_counter()
            restate.forward(sd, *[f[k] for k in ARGS], mode, run_dead=True)
            dt = time.perf_counter() - t
            if i > 0:  # first frame = warm-up
                times.append(dt)
    return times


def torch_eager_gpu_baseline(sd, num_pc, n_frames, dev, mode="val"):
    """Secondary baseline of the cpu_baseline leg: the same oracle port executed by stock PyTorch
    on the B200 (ATen / cuDNN / cuBLAS eager kernels, one frame per forward as the reference runs) -- what the
    reference's own code gets from this GPU without this library.  Returns seconds per frame."""
    from cofii2p_b200.frames import frame_to, make_frame
    from oracle import restate
    gsd = {k: v.to(dev) for k, v in sd.items()}
    frames = [frame_to(make_frame(100 + i, num_pc=num_pc, cache_dir="/tmp/cofi_frames", device=str(dev)), dev)
              for i in range(2)]
    times = []
    with torch.no_grad():
        for i in range(n_frames + 2):
            f = frames[i % 2]
            torch.cuda.synchronize(dev)
            t = time.perf_counter()
            restate.forward(gsd, *[f[k] for k in ARGS], mode, run_dead=True)
            torch.cuda.synchronize(dev)
            if i >= 2:  # two warm-ups (cuDNN autotune, allocator)
                times.append(time.perf_counter() - t)
    return times


# ------------------------------------------------------------------------------------------------ arms
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cofii2p_b200.model.network import CoFiI2P
    from cofii2p_b200.options import Options_KITTI
    from cofii2p_b200.weights import seeded_state_dict
    from cofii2p_b200.frames import make_frame
    from oracle import restate
    m = CoFiI2P(Options_KITTI())
    sd = seeded_state_dict(m, 0)
    cores, cands = pick_cpu_threads(sd)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    frames = [make_frame(100 + i, num_pc=args.num_pc, cache_dir="/tmp/cofi_frames", device=dev)
              for i in range(min(args.steps + args.warmup, 4))]
    with torch.no_grad():
        for i in range(args.warmup):
            f = frames[i % len(frames)]
            restate.forward(sd, *[f[k] for k in ARGS], "val", run_dead=True)
        t0 = time.perf_counter()
        for i in range(args.steps):
            f = frames[(args.warmup + i) % len(frames)]
            restate.forward(sd, *[f[k] for k in ARGS], "val", run_dead=True)
        dt = time.perf_counter() - t0
    fps = args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "reference CoFiI2P.forward(val) on CPU, one 20480-pt frame per step "
                               "(bounded sample of configs[1])", "num_pc": args.num_pc, "frames_per_step": 1},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} frames, oracle/restate.py forward incl. dead layer3/4, "
                                   f"torch {torch.__version__} CPU, best of thread counts {cands} on a "
                                   f"{os.cpu_count()}-core host -> {torch.get_num_threads()} threads"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_cofi(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from cofii2p_b200 import lib, ops
    from cofii2p_b200.engine import InferenceEngine
    from cofii2p_b200.frames import make_frame, stack_frames

    ops.set_engine(args.engine)
    model, sd = build_model(dev)
    B = args.batch
    frames = [make_frame(rank * B + i, num_pc=args.num_pc, cache_dir="/tmp/cofi_frames", device=f"cuda:{local}")
              for i in range(B)]
    batch = stack_frames(frames)
    eng = InferenceEngine(model, batch, mode="val", use_graph=not args.no_graph)
    host = eng.host_buffers(batch)
    stream = eng.stream

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def timed_on(st, fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with torch.cuda.stream(st):
            e0.record(st)
            for _ in range(steps):
                fn()
            e1.record(st)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident throughput (inputs already in HBM) -------------------------------------------
    for _ in range(max(args.warmup, 3)):
        eng.run()
    clocks = ClockSampler(local)
    ms_total = timed(eng.run, args.steps)
    clk = clocks.stop()
    frames_total = world * B * args.steps
    value = frames_total / (ms_total / 1000.0)

    # ---- end to end: pinned host buffers -> H2D -> forward -> D2H (double-buffered public API) ---------
    from cofii2p_b200.engine import PipelinedEngine
    pipe = PipelinedEngine(model, batch, depth=2)
    hosts = [host, eng.host_buffers(batch)]  # two pinned staging copies, as a loader with prefetch would own
    io = {"in": 0, "out": 0}
    for i in range(3):
        pipe.step(hosts[i % 2])
    pipe.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    e0.record(pipe.h2d)
    for i in range(args.steps):
        io["in"], io["out"] = pipe.step(hosts[i % 2])
    pipe.compute.wait_stream(pipe.h2d)
    pipe.d2h.wait_stream(pipe.compute)
    e1.record(pipe.d2h)
    pipe.synchronize()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    barrier()
    ms_e2e_t = torch.tensor([max(e0.elapsed_time(e1), 0.0)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_e2e_t, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms_e2e_t.item())
    e2e_value = frames_total / (ms_e2e / 1000.0)
    pipe.last_results()  # validates the err flag / keeps the API honest
    del pipe

    # ---- the same pipeline fed with the point pyramid only: the 13 KNN-128 tables are built on the device inside the
    # graph (csrc/knn.cu) instead of being computed on the host and shipped over PCIe -------
    from_points = None
    if not args.no_from_points:
        eng_p = InferenceEngine(model, batch, mode="val", use_graph=not args.no_graph, tables="device")
        for _ in range(3):
            eng_p.run()
        ms_p = timed_on(eng_p.stream, eng_p.run, args.steps)
        lp = eng_p.launches_per_step
        del eng_p
        pipe = PipelinedEngine(model, batch, depth=2, tables="device")
        hosts_p = [pipe.engines[0].host_buffers(batch), pipe.engines[0].host_buffers(batch)]
        for i in range(3):
            pipe.step(hosts_p[i % 2])
        pipe.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(pipe.h2d)
        for i in range(args.steps):
            pin, pout = pipe.step(hosts_p[i % 2])
        pipe.compute.wait_stream(pipe.h2d)
        pipe.d2h.wait_stream(pipe.compute)
        e1.record(pipe.d2h)
        pipe.synchronize()
        barrier()
        ms_pe = torch.tensor([max(e0.elapsed_time(e1), 0.0)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms_pe, op=dist.ReduceOp.MAX)
        from_points = {"value": frames_total / (ms_p / 1000.0), "e2e": frames_total / (float(ms_pe.item()) / 1000.0),
                       "unit": UNIT, "h2d_bytes_per_step": pin, "d2h_bytes_per_step": pout, "launches_per_step": lp,
                       "what": "same forward, inputs = point pyramid + image only; the KNN-128 index tables (neighbors, "
                               "subsampling: 128 columns; upsampling: its single live column) are built by cofi_knn_pyramid "
                               "inside the captured graph"}
        pipe.last_results()
        del pipe

    # ---- roofline of the dominant kernel family: eager pass bracketed by CUDA events per launch --------
    hbm, tf_burst, tf_sust, peaks_src = measured_peaks()
    model.fork_image_stream = False  # serialise the two branches so per-kernel event times do not overlap
    with torch.no_grad(), torch.cuda.stream(stream):
        eng._step_eager()
        torch.cuda.synchronize(dev)
        ops.profile_start()
        for _ in range(2):
            eng._step_eager()
        # contractions are split per call at the ridge of the tf32 tensor roof (half the measured bf16 rate) and the HBM roof
        prof = ops.profile_stop(ridge=0.5 * tf_sust * 1e12 / (hbm * 1e9))
    model.fork_image_stream = True
    # kernel families: every tcgen05 GEMM entry point (plain / +column statistics / fp16 operands / +LayerNorm) is the
    # same kernel template (gemm_tc_kernel); the two KPConv aggregate variants likewise
    fam = {}
    for name, d in prof.items():
        base, _, bound = name.partition("|")
        key = "cofi_gemm*" if base.startswith("cofi_gemm") else ("cofi_kpconv_aggregate*" if base.startswith("cofi_kpconv_aggregate") else base)
        if bound:
            key += " [" + bound + "-bound calls]"
        f = fam.setdefault(key, dict(calls=0, ms=0.0, flops=0.0, bytes=0.0))
        for k in f:
            f[k] += d[k]
    tot_ms = sum(d["ms"] for d in fam.values())
    name, d = max(fam.items(), key=lambda kv: kv[1]["ms"])
    tensor_ops = ("cofi_gemm* [tensor-bound calls]", "cofi_conv2d_nhwc [tensor-bound calls]", "cofi_attention_vt", "cofi_attention",
                  "cofi_sim_argmin")
    if name in tensor_ops:
        ach = d["flops"] / (d["ms"] / 1e3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": tf_sust, "unit": "TFLOP/s", "frac": ach / tf_sust}
    else:
        ach = d["bytes"] / (d["ms"] / 1e3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm}
    # the family mixes tensor-bound (large K) and HBM-bound (K <= 128) launches: report the other roof too
    roof["hbm_gbs_same_family"] = d["bytes"] / (d["ms"] / 1e3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(name.split(" [")[0])   # ncu average over ALL launches of the kernel family
    roof.update({"traffic": traffic, "kernel": name, "launches_profiled": d["calls"],
                 "avg_launch_us": 1000.0 * d["ms"] / d["calls"], "share_of_step": d["ms"] / tot_ms,
                 "peak_source": peaks_src + (" (bf16 dense sustained; the family runs tf32 (nominal peak = half of bf16) and "
                                             "fp16 operands)" if name in tensor_ops else " (copy bandwidth)"),
                 "by_kernel_ms_per_step": {k: round(v["ms"] / 2, 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}})

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "tf32": "tf32", "tf32x3": "tf32x3"}[args.engine], "data": "synthetic",
        "config": {"workload": "configs[1]: single-GPU inference, batch=8 synthetic KITTI frames (3x160x512 img, "
                               "20480x3 cloud, 5-level KNN-128 tables), val-style forward",
                   "frames_per_gpu_per_step": B, "num_pc": args.num_pc, "engine": args.engine,
                   "cuda_graph": not args.no_graph, "parallelism": f"replicas x{world}",
                   "l2": "inputs larger than L2 (index tables 0.49 GB per step vs 126 MB L2)"},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": io["in"], "d2h_bytes_per_step": io["out"],
                "ms_per_step": ms_e2e / args.steps, "wall_ms_per_step": wall_ms / args.steps,
                "api": "cofii2p_b200.engine.PipelinedEngine.step(pinned host batch): H2D || graph replay || D2H"},
        "gpu_launches": eng.launches_per_step * args.steps,
        "launches_per_step": eng.launches_per_step,
        "roofline": roof,
    }
    if from_points is not None:
        line["from_points"] = from_points
    if world == 1 and not args.no_cpu_baseline:
        cpu_sd = {k: v.detach().cpu() for k, v in sd.items()}
        cores, cands = pick_cpu_threads(cpu_sd)
        times = cpu_forward_baseline(cpu_sd, args.num_pc, args.cpu_frames)
        fps = len(times) / sum(times)
        line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{len(times)} frames (after 1 warm-up) of the same 20480-pt workload, "
                                          f"oracle/restate.py forward(val) incl. dead layer3/4; best of thread counts "
                                          f"{cands} on a {os.cpu_count()}-core host -> {cores} threads"}
        try:
            tg = torch_eager_gpu_baseline(cpu_sd, args.num_pc, 5, dev)
            line["cpu_baseline"]["torch_eager_b200"] = {
                "value": len(tg) / sum(tg), "unit": UNIT,
                "sample": f"{len(tg)} frames after 2 warm-ups: the same oracle port run by stock PyTorch eager kernels "
                          "(ATen/cuDNN/cuBLAS, fp32) on this B200, one frame per forward"}
        except Exception as e:  # a baseline must never take the measurement down
            line["cpu_baseline"]["torch_eager_b200"] = {"unavailable": repr(e)[:200]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.workload == "train":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import train_bench
        (train_bench.run_reference if args.impl == "reference" else train_bench.run_train)(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_cofi(args)


if __name__ == "__main__":
    main()
