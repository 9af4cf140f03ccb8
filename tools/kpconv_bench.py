"""BASELINE config 3: KPConv encoder (KPConvFPN: 14 KPConv layers + 34 linears + 3-stage decoder) standalone,
N0 sweep incl. the named 40960-point cloud.  Reports time, effective (logical) gather GB/s, compulsory-bytes roofline
 and per-kernel-family times, one frame per run, CUDA-graph replay timed with CUDA events."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_model, measured_peaks
from cofii2p_b200 import ops
from cofii2p_b200.frames import make_frame, frame_to

ops.set_engine(sys.argv[1] if len(sys.argv) > 1 else "tf32")
dev = torch.device("cuda", 0)
model, _ = build_model(dev)
enc = model.pc_encoder
hbm, *_ = measured_peaks()
LAYERS = [(1, 4, 64), (1, 32, 32), (2, 32, 32), (2, 64, 64), (2, 64, 64), (4, 64, 64), (4, 128, 128), (4, 128, 128),
          (8, 128, 128), (8, 256, 256), (8, 256, 256), (16, 256, 256), (16, 512, 512), (16, 512, 512)]  # (N0/M, C, Cout)
for n0 in (10240, 20480, 40960, 81920):
    f = frame_to(make_frame(7, num_pc=n0, cache_dir="/tmp/cofi_frames", device="cuda"), dev)
    d = f["pc_data_dict"]
    with torch.no_grad():
        for _ in range(2):
            enc(d, 1)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = enc(d, 1)
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ops.profile_start(); enc(d, 1); prof = ops.profile_stop()
    ms = sorted(ts)[len(ts) // 2]
    comp = logical = 0.0
    for div, c, co in LAYERS:
        m = n0 // div
        n = m * 2 if div in (2, 4, 8, 16) and (div, c, co) in ((2, 32, 32), (4, 64, 64), (8, 128, 128), (16, 256, 256)) else m
        comp += 8 * m * 128 + 4 * n * c + 12 * n + 12 * m + 4 * m * co + 4 * 15 * c * co
        logical += m * 128 * (4 * c + 12 + 8) + 4 * m * co
    agg = sum(v["ms"] for k, v in prof.items() if "kpconv_aggregate" in k)
    rec = dict(n0=n0, engine=ops.get_engine(), encoder_ms=ms, frames_per_s=1000.0 / ms,
               kpconv_aggregate_ms=agg, compulsory_GB=comp / 1e9, logical_gather_GB=logical / 1e9,
               effective_gather_GBps_aggregate=logical / 1e9 / (agg / 1e3),
               compulsory_GBps_aggregate=comp / 1e9 / (agg / 1e3), hbm_peak_GBps=hbm,
               by_family_ms={k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8]})
    print(json.dumps(rec))
