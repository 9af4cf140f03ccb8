"""Does the row order of a level matter to the gather kernels?  KPConv aggregate, neighbour max-pool and the whole point
encoder on one 8-frame batch with the generator's shuffled point order vs the same frames re-labelled in Morton order
(points, features and all index tables permuted consistently; the results are the same rows in another order)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_model
from cofii2p_b200 import ops
from cofii2p_b200.frames import make_frame, stack_frames
from reorder import morton_permutations, permute_pyramid

ops.set_engine("tf32")
dev = torch.device("cuda", 0)
model, _ = build_model(dev)
B = 8
batch = stack_frames([make_frame(i, cache_dir="/tmp/cofi_frames", device="cuda") for i in range(B)])
d0 = {k: ([t.to(dev) for t in v] if isinstance(v, list) and torch.is_tensor(v[0]) else (v.to(dev) if torch.is_tensor(v) else v))
      for k, v in batch["pc_data_dict"].items()}
perms = morton_permutations(d0["points"], B)
d1 = permute_pyramid(d0, perms, B)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


res = {}
for name, d in (("shuffled", d0), ("morton", d1)):
    with torch.no_grad():
        x64 = torch.randn(d["points"][0].shape[0], 64, device=dev)
        xh = ops.cast_f16(x64)
        packed = ops.pack_points(d["points"][0], x64)
        kp = model.pc_encoder.encoder1_2.KPConv
        r = {}
        r["maxpool_f16_l0_to_l1_us"] = timed(lambda: ops.maxpool_rows_f16(xh, d["subsampling"][0], B))
        x32 = torch.randn(d["points"][0].shape[0], 32, device=dev)
        packed32 = ops.pack_points(d["points"][0], x32)
        r["aggregate_f16_l0_C32_us"] = timed(lambda: ops.kpconv_aggregate_f16(x32, packed32, d["points"][0], d["neighbors"][0],
                                                                              kp.kernel_points, kp.sigma, B, kp.kp_reach()))
        x128 = torch.randn(d["points"][2].shape[0], 128, device=dev)
        packed128 = ops.pack_points(d["points"][2], x128)
        kp3 = model.pc_encoder.encoder3_2.KPConv
        r["aggregate_f16_l2_C128_us"] = timed(lambda: ops.kpconv_aggregate_f16(x128, packed128, d["points"][2], d["neighbors"][2],
                                                                               kp3.kernel_points, kp3.sigma, B, kp3.kp_reach()))
        r["encoder_ms"] = timed(lambda: model.pc_encoder(d, B), 5) / 1e3
    res[name] = r
print(json.dumps(res))
