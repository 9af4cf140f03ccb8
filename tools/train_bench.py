#!/usr/bin/env python
"""Training-step benchmark (BASELINE.json configs[4]): data-parallel training, batch = 4 synthetic KITTI frames per GPU,
the reference's three losses (model/loss.py), NCCL gradient all-reduce, Adam -- at 1/2/4/8 x B200.

    python bench.py --workload train --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --workload train --gpus N --steps K --warmup W
    python bench.py --workload train --impl reference --steps 1 --warmup 0      # CPU arm (oracle autograd)

One step = forward (train mode) + losses + backward + all-reduce + Adam over the rank's 4 stacked frames.  value = frames/s
over all ranks with the batch resident in HBM; e2e = the same with the batch coming from pinned host memory every step
and the loss read back.  Same JSON line as bench.py."""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "training frames/sec (20480 pts, 160x512 img, fwd+loss+bwd+allreduce+Adam)"
UNIT = "frames/s"
SUP = ("pc_kpt_idx", "pc_outline_idx", "coarse_img_kpt_idx", "K_4", "P", "fine_xy", "fine_center_kpt_coors")


def _map(x, fn):
    if torch.is_tensor(x):
        return fn(x)
    if isinstance(x, list):
        return [_map(y, fn) for y in x]
    if isinstance(x, dict):
        return {k: _map(v, fn) for k, v in x.items()}
    return x


def _nbytes(x):
    if torch.is_tensor(x):
        return x.numel() * x.element_size()
    if isinstance(x, (list, tuple)):
        return sum(_nbytes(y) for y in x)
    if isinstance(x, dict):
        return sum(_nbytes(v) for v in x.values())
    return 0


def cpu_train_step(sd, frame, opt, threads=None):
    """The reference's training iteration on the host: oracle forward (train mode) + losses + autograd backward."""
    from cofii2p_b200.train import training_losses
    from oracle import restate
    if threads:
        torch.set_num_threads(threads)
    sdr = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k and "kernel_points" not in k
               else v.clone()) for k, v in sd.items()}
    t = time.perf_counter()
    out = restate.forward(sdr, frame["pc_data_dict"], frame["img"], frame["fine_center_kpt_coors"], frame["fine_xy"],
                          frame["fine_pc_inline_index"], "train", run_dead=True, bn_training=True)
    loss, _, _ = training_losses(out, {k: frame[k] for k in SUP}, opt, frame["pc_data_dict"]["points"][-1])
    loss.backward()
    return time.perf_counter() - t, float(loss.detach())


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from cofii2p_b200.frames import make_frame
    from cofii2p_b200.model.network import CoFiI2P
    from cofii2p_b200.options import Options_KITTI
    from cofii2p_b200.weights import seeded_state_dict
    opt = Options_KITTI()
    sd = seeded_state_dict(CoFiI2P(opt), 0)
    cores = min(16, os.cpu_count() or 1)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    frames = [make_frame(100 + i, num_pc=args.num_pc, cache_dir="/tmp/cofi_frames", device=dev) for i in range(2)]
    for i in range(args.warmup):
        cpu_train_step(sd, frames[i % 2], opt, cores)
    t0 = time.perf_counter()
    for i in range(args.steps):
        cpu_train_step(sd, frames[i % 2], opt, cores)
    dt = time.perf_counter() - t0
    fps = args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "reference training iteration on CPU (forward train mode + 3 losses + backward), one 20480-pt "
                               "frame per step (bounded sample of configs[4])", "num_pc": args.num_pc, "frames_per_step": 1},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} iterations, oracle/restate.py forward + torch autograd backward, {cores} threads"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)


def train_leg(args, dev, world, rank, local, steps=None):
    """The data-parallel training step (configs[4]) as one leg of bench.py's default line: 4 frames per GPU, forward (train
    mode) + the reference's three losses + backward captured in a CUDA graph, NCCL gradient all-reduce over NVLink, fused
    Adam.  Returns a dict for rank 0's JSON line: frames/s over all ranks (max over ranks of the device time), ms per step,
    and the all-reduce cost: `allreduce_ms_isolated` (the collective timed alone) and `allreduce_ms_exposed` (step time with
    the collective minus the same step with it skipped) -- all measured in this run."""
    import torch.distributed as dist
    import bench
    from cofii2p_b200 import ops
    from cofii2p_b200.frames import make_frame, stack_frames
    from cofii2p_b200.options import Options_KITTI
    from cofii2p_b200.shard import frames_for_rank
    from cofii2p_b200.train import TrainStep
    steps = steps or args.steps
    engine = "tf32" if args.engine == "parity" else args.engine   # training runs the tf32 contraction engine
    ops.set_engine(engine)
    model, _ = bench.build_model(dev)
    model.train()
    opt = Options_KITTI()
    B = args.train_batch
    frames = [make_frame(s, num_pc=args.num_pc, cache_dir="/tmp/cofi_frames", device=f"cuda:{local}")
              for s in frames_for_rank(rank, world, B)]
    batch = _map(stack_frames(frames), lambda t: t.to(dev))
    ts = TrainStep(model, opt)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    losses = []
    for _ in range(2):
        losses.append(float(ts.step(batch)[0]))
    if not args.no_graph:
        ts.enable_cuda_graph(batch)
    for _ in range(3):
        ts.step(batch)
    ms = timed(lambda: ts.step(batch), steps)
    losses.append(float(ts.step(batch)[0]))
    ts.check_errors()
    out = {"metric": METRIC, "value": world * B * steps / (ms / 1e3), "unit": UNIT, "ms_per_step": ms / steps,
           "frames_per_gpu_per_step": B, "engine": engine, "cuda_graph": ts.graph is not None, "parallelism": f"dp{world}",
           "live_parameters": int(ts.flat_p.numel()), "loss_first_last": [losses[0], losses[-1]],
           "allreduce_bytes_per_step": int(ts.flat_g.numel() * 4) if world > 1 else 0,
           "efficiency_basis": "weak scaling: 4 frames per GPU at every N; efficiency = value(N) / (N * value(1)) from the "
                               "train.value of the N=1 line of the same scaling run",
           "what": "configs[4]: forward(train) + desc/overlap/fine circle losses + backward (one CUDA graph), one NCCL all-reduce "
                   "of the flat gradient, one fused Adam kernel"}
    if world > 1:
        ms_ar = timed(lambda: dist.all_reduce(ts.flat_g), steps)
        w, ts.world = ts.world, 1          # the same step with the collective skipped (timing only; gradients unscaled)
        ms_no = timed(lambda: ts.step(batch), steps)
        ts.world = w
        out["allreduce_ms_isolated"] = ms_ar / steps
        out["allreduce_ms_exposed"] = max(ms - ms_no, 0.0) / steps
        out["ms_per_step_without_allreduce"] = ms_no / steps
    del ts, model
    torch.cuda.empty_cache()
    return out


def run_train(args):
    import torch.distributed as dist
    import bench
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from cofii2p_b200 import lib, ops
    from cofii2p_b200.frames import make_frame, stack_frames
    from cofii2p_b200.options import Options_KITTI
    from cofii2p_b200.shard import frames_for_rank
    from cofii2p_b200.train import TrainStep

    ops.set_engine(args.engine)
    model, sd = bench.build_model(dev)
    model.train()
    opt = Options_KITTI()
    B = args.train_batch
    frames = [make_frame(s, num_pc=args.num_pc, cache_dir="/tmp/cofi_frames", device=f"cuda:{local}")
              for s in frames_for_rank(rank, world, B)]
    host = _map(stack_frames(frames), lambda t: t.pin_memory())
    batch = _map(host, lambda t: t.to(dev))
    ts = TrainStep(model, opt)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    losses = []
    per_step = 0
    for _ in range(max(args.warmup, 3)):
        n0 = lib.launch_count()
        losses.append(ts.step(batch)[0])
        per_step = lib.launch_count() - n0
    if not args.no_graph:
        ts.enable_cuda_graph(batch)      # forward + losses + backward replayed as one CUDA graph; all-reduce + Adam eager
        for _ in range(2):
            losses.append(ts.step(batch)[0].clone())
    clocks = bench.ClockSampler(local)
    ms_total = timed(lambda: losses.append(ts.step(batch)[0].clone()), args.steps)
    launches = per_step * args.steps
    clk = clocks.stop()
    ts.check_errors()
    frames_total = world * B * args.steps
    value = frames_total / (ms_total / 1e3)

    # ---- end to end: pinned host batch -> H2D, step, loss back to the host, every step ---------------------------
    h2d, d2h = _nbytes(host), 4

    def e2e_step():
        # graph mode: the pinned host batch is copied straight into the captured step's input buffers
        b = host if ts.graph is not None else _map(host, lambda t: t.to(dev, non_blocking=True))
        loss, _ = ts.step(b)
        losses.append(float(loss.item()))

    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = frames_total / (ms_e2e / 1e3)

    # ---- per-kernel split of one step (CUDA events per launch) ------------------------------------------------------
    hbm, tf_burst, tf_sust, peaks_src = bench.measured_peaks()
    model.fork_image_stream = False
    graph, ts.graph = ts.graph, None     # the per-kernel split needs an eager step
    ops.profile_start()
    ts.step(batch)
    prof = ops.profile_stop()
    ts.graph = graph
    fam = {}
    for name, d in prof.items():
        key = "cofi_gemm*" if name.startswith("cofi_gemm") and not name.startswith("cofi_gemm_tn") else name
        f = fam.setdefault(key, dict(calls=0, ms=0.0, flops=0.0, bytes=0.0))
        for k in f:
            f[k] += d[k]
    tot_ms = sum(d["ms"] for d in fam.values())
    name, d = max(fam.items(), key=lambda kv: kv[1]["ms"])
    if d["flops"] > 0 and d["flops"] / max(d["bytes"], 1.0) > tf_sust * 1e12 / (hbm * 1e9) / 8:
        ach = d["flops"] / (d["ms"] / 1e3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": tf_sust, "unit": "TFLOP/s", "frac": ach / tf_sust}
    else:
        ach = d["bytes"] / (d["ms"] / 1e3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm}
    roof.update({"traffic": None, "kernel": name, "launches_profiled": d["calls"], "avg_launch_us": 1e3 * d["ms"] / d["calls"],
                 "share_of_step": d["ms"] / tot_ms, "peak_source": peaks_src,
                 "kernel_ms_sum": tot_ms,
                 "by_kernel_ms_per_step": {k: round(v["ms"], 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}})
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    lf = [float(x) for x in losses]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "tf32": "tf32", "tf32x3": "tf32x3"}[args.engine], "data": "synthetic",
        "config": {"workload": "configs[4]: data-parallel training step, batch=4/GPU synthetic KITTI frames, circle losses, "
                               "NCCL gradient all-reduce, Adam", "frames_per_gpu_per_step": B, "num_pc": args.num_pc,
                   "engine": args.engine, "cuda_graph": ts.graph is not None, "parallelism": f"dp{world}", "live_parameters": int(ts.flat_p.numel()),
                   "allreduce_bytes_per_step": int(ts.flat_g.numel() * 4) if world > 1 else 0,
                   "l2": "inputs and saved activations far larger than L2 (126 MB)"},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps, "api": "cofii2p_b200.train.TrainStep.step(batch from pinned host memory) + loss.item()"},
        "gpu_launches": launches, "launches_per_step": launches // max(args.steps, 1),
        "loss_first_last": [lf[0], lf[-1]],
        "roofline": roof,
    }
    if world == 1 and not args.no_cpu_baseline:
        cpu_sd = {k: v.detach().cpu() for k, v in sd.items()}
        cores = min(16, os.cpu_count() or 1)
        fr = make_frame(100, num_pc=args.num_pc, cache_dir="/tmp/cofi_frames", device=f"cuda:{local}")
        dt, _ = cpu_train_step(cpu_sd, fr, opt, cores)
        line["cpu_baseline"] = {"value": 1.0 / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "1 training iteration (forward train mode incl. dead layers + losses + autograd backward) of "
                                          "one 20480-pt frame, oracle/restate.py on the host"}
        try:   # secondary baseline: the same iteration by stock PyTorch eager + autograd on this B200
            from cofii2p_b200.frames import frame_to
            gsd = {k: v.to(dev) for k, v in cpu_sd.items()}
            gfr = frame_to(fr, dev)
            ts = []
            for i in range(5):
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                cpu_train_step(gsd, gfr, opt, None)
                torch.cuda.synchronize(dev)
                if i >= 2:
                    ts.append(time.perf_counter() - t0)
            line["cpu_baseline"]["torch_eager_b200"] = {
                "value": len(ts) / sum(ts), "unit": UNIT,
                "sample": f"{len(ts)} iterations after 2 warm-ups: oracle forward + losses + torch autograd backward run by "
                          "stock PyTorch eager kernels (fp32) on this B200, one frame per iteration, no optimizer step"}
        except Exception as e:
            line["cpu_baseline"]["torch_eager_b200"] = {"unavailable": repr(e)[:200]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
