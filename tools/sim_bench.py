"""Fused similarity + arg-min throughput (BASELINE sweep point 10240 x 20480 x 64 and the coarse 1280 x 1280 x 128 case).

Per shape and engine: median of 7 launches (L2 flushed in between), CUDA events on the launching stream, clocks sampled
during the run.  Engines: `exact-tc` = the product path (tcgen05 fp16 candidate pass + exact fp32 re-rank, bit-identical
to the fp32 engine; both launches timed together, fp16 copies resident as l2norm_rows_f16 leaves them), `tcgen05-f16` = the
candidate pass's plain arg-min variant, `tcgen05-tf32`, `simt-fp32`.  `frac_*`: against MEASURED_PEAKS.json (bf16 dense,
burst -- the kernel is timed alone) and against the TMEM-read bound 148 SM x 64 B/clk x clock (every score has to be read
from TMEM once; /opt/skills/guides/B300_MICROARCH.md "LDTM throughput").   usage: python tools/sim_bench.py > out.jsonl"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

import bench
from cofii2p_b200 import ops

hbm, tf_burst, tf_sust, src = bench.measured_peaks()
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
clocks = bench.ClockSampler(0)
recs = []
for (npt, npx, c, frames) in [(10240, 20480, 64, 1), (10240, 20480, 64, 8), (1280, 1280, 128, 8), (20480, 20480, 128, 4),
                              (10240, 20480, 128, 8)]:
    pt = F.normalize(torch.randn(frames * npt, c, device="cuda"), dim=1)
    px = F.normalize(torch.randn(frames * npx, c, device="cuda"), dim=1)
    ptn, pth = ops.l2norm_rows_f16(pt)
    pxn, pxh = ops.l2norm_rows_f16(px)
    stats = torch.zeros(2, dtype=torch.int32, device="cuda")
    for eng, name in (("exact", "exact-tc"), (None, "tcgen05-f16"), (ops.ENGINE_TF32, "tcgen05-tf32"), (ops.ENGINE_FP32, "simt-fp32")):
        if name == "simt-fp32" and frames * npt * npx > 4e9:
            continue
        if eng == "exact":
            run = lambda: ops.sim_argmin(ptn, pxn, frames, pt_h=pth, px_h=pxh, stats=stats)
        elif eng is None:
            run = lambda: ops.sim_argmin_f16(pth, pxh, frames)
        else:
            run = lambda: ops.sim_argmin(ptn, pxn, frames, engine=eng)
        for _ in range(2):
            run()
        stats.zero_()
        ts = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        fl = 2.0 * frames * npt * npx * c
        rec = dict(engine=name, npt=npt, npx=npx, c=c, frames=frames, us=t * 1e3, tflops=fl / t / 1e9,
                   frac_of_bf16_burst=fl / t / 1e9 / tf_burst, matrix_bytes_never_written=4.0 * frames * npt * npx)
        if eng == "exact":
            s = stats.tolist()
            rec["reranked_candidates_per_point"] = s[0] / (7.0 * frames * npt)
            rec["full_scan_rows"] = s[1]
        recs.append(rec)
clk = clocks.stop()
mhz = clk.get("sm_mhz") or 1965.0
for rec in recs:
    tmem_roof = 148 * 16 * mhz * 1e6 * 2 * rec["c"] / 1e12      # scores/s readable from TMEM x 2C flop per score
    rec["tmem_read_roof_tflops"] = tmem_roof
    rec["frac_of_tmem_read_roof"] = rec["tflops"] / tmem_roof
    rec["clocks"] = clk
    rec["peak_bf16_burst_tflops"] = tf_burst
    rec["peak_source"] = src
    print(json.dumps(rec))
