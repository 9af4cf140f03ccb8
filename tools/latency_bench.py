"""Single-frame latency / FPS of the drop-in `CoFiI2P.forward` (the reference's own efficiency methodology,
evaluation/get_flops_fps.py:14-63: warm-up, then a timed loop bracketed by cuda.synchronize), test and val modes,
eager launches vs the cached CUDA graph."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_model
from cofii2p_b200 import ops
from cofii2p_b200.frames import make_frame, frame_to

ops.set_engine(sys.argv[1] if len(sys.argv) > 1 else "tf32")
dev = torch.device("cuda", 0)
model, _ = build_model(dev)
frames = [frame_to(make_frame(i, cache_dir="/tmp/cofi_frames", device="cuda"), dev) for i in range(4)]
A = ("pc_data_dict", "img", "fine_center_kpt_coors", "fine_xy", "fine_pc_inline_index")
for graph in (False, True):
    model.enable_cuda_graph(graph)
    for mode in ("val", "test"):
        with torch.no_grad():
            for i in range(5):
                model(*[frames[i % 4][k] for k in A], mode)
            torch.cuda.synchronize()
            n = 50
            t = time.perf_counter()
            for i in range(n):
                model(*[frames[i % 4][k] for k in A], mode)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t) / n
        print(json.dumps(dict(engine=ops.get_engine(), cuda_graph=graph, mode=mode, latency_ms=dt * 1e3, fps=1.0 / dt)))
