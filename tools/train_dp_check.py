#!/usr/bin/env python
"""Data-parallel training parity on N GPUs (run under torchrun): after the NCCL all-reduce every rank holds the same
gradient, it equals the gradient of the union batch computed by one process, and the parameters stay identical across
ranks after Adam steps.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/train_dp_check.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from cofii2p_b200 import ops
    from cofii2p_b200.frames import frame_to, make_frame, stack_frames
    from cofii2p_b200.model.network import CoFiI2P
    from cofii2p_b200.options import Options_KITTI
    from cofii2p_b200.shard import frames_for_rank
    from cofii2p_b200.train import TrainStep
    from cofii2p_b200.weights import seeded_state_dict
    ops.set_engine("fp32")
    opt = Options_KITTI()

    def fresh():
        m = CoFiI2P(opt)
        m.load_state_dict(seeded_state_dict(m, 0), strict=True)
        return m.to(dev)

    B = 2
    mine = [frame_to(make_frame(s, num_pc=4096, cache_dir="/tmp/cofi_frames", device=f"cuda:{local}"), dev)
            for s in frames_for_rank(rank, world, B)]
    ts = TrainStep(fresh(), opt)
    ts.step(stack_frames(mine))                      # backward + all-reduce + Adam
    g_avg = ts.flat_g.clone() / world                 # what Adam consumed
    # (1) identical gradient on every rank
    ref = g_avg.clone()
    dist.broadcast(ref, 0)
    same_grad = bool(torch.equal(ref, g_avg))
    # (2) equals the union batch on one process
    everyone = [frame_to(make_frame(s, num_pc=4096, cache_dir="/tmp/cofi_frames", device=f"cuda:{local}"), dev)
                for r in range(world) for s in frames_for_rank(r, world, B)]
    solo = TrainStep(fresh(), opt)
    solo.world = 1
    solo.backward(stack_frames(everyone))
    rel = float((solo.flat_g - g_avg).norm() / solo.flat_g.norm())
    # (3) parameters stay in lock-step over further steps (graph replay included)
    ts.enable_cuda_graph(stack_frames(mine))
    for _ in range(3):
        ts.step(stack_frames(mine))
    p = ts.flat_p.clone()
    p0 = p.clone()
    dist.broadcast(p0, 0)
    same_param = bool(torch.equal(p0, p))
    flags = torch.tensor([int(same_grad), int(same_param), int(rel < 1e-4)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "frames_per_rank": B, "same_gradient_on_all_ranks": bool(flags[0]),
                          "same_parameters_after_4_steps": bool(flags[1]), "union_batch_rel_err": rel,
                          "union_batch_ok": bool(flags[2])}))
    dist.destroy_process_group()
    assert bool(flags.min())


if __name__ == "__main__":
    main()
