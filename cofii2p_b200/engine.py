"""Batched inference engine: static device buffers + one CUDA graph for the whole forward, `val` or `test` mode.

The reference runs one frame at a time through ~1,150 eager kernels and 256 host syncs.  Here B
frames are stacked along the row axis (frame-local index tables), every kernel launch of the forward is
captured once into a CUDA graph, and a step is: (optional) async H2D copies from pinned host memory into the
static input buffers -> graph replay -> (optional) async D2H of the outputs into pinned host memory.
Only column 0 of the up-sampling tables is ever read by the model (reference model/kpconv/functional.py:20), so
the engine keeps and uploads [N,1] columns instead of [N,128] tables.

mode="test" is what the reference's evaluation runs (evaluation/eval_all.py:96): the whole matching stage -- fused
similarity + arg-min over all B frames (tcgen05 candidate pass + exact fp32 re-rank), threshold loop + border mask +
compaction, point2node, patch / feature gathers, and the caller's 16-way fine match (eval_all.py:99-102) -- is inside the
captured graph with fixed shapes (N4 rows per frame, the upper bound on the matches) and the match counts in a device
tensor; nothing synchronises with the host until results() trims the rows.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import lib as _libmod
from . import ops


def _pin(t: torch.Tensor) -> torch.Tensor:
    return t.contiguous().pin_memory()


class InferenceEngine:
    def __init__(self, model, batch: Dict, mode: str = "val", use_graph: bool = True, stream=None, tables: str = "host",
                 knn_stream=None):
        """`batch`: output of frames.stack_frames (CPU or CUDA tensors) that fixes all shapes.
        `stream`: compute stream shared by several engines (PipelinedEngine).
        `tables`: "host" = the KNN index tables arrive with the batch, as the reference's data loader supplies them;
        "device" = only the point pyramid arrives and the tables are built inside the captured graph by
        ops.knn_pyramid (csrc/knn.cu; neighbours/subsampling k columns, upsampling its single live column)."""
        assert mode in ("val", "test")
        assert tables in ("host", "device")
        self.model, self.mode, self.B, self.tables = model, mode, batch["frames"], tables
        dev = next(model.parameters()).device
        self.device = dev
        d = batch["pc_data_dict"]
        self.inp = {
            "points": [t.to(dev).contiguous() for t in d["points"]],
            "neighbors": [t.to(dev).contiguous() for t in d["neighbors"]],
            "subsampling": [t.to(dev).contiguous() for t in d["subsampling"]],
            "upsampling": [t[:, :1].to(dev).contiguous() for t in d["upsampling"]],
            "feats": d["feats"].to(dev).contiguous(),
            "lengths": d["lengths"],
        }
        self.knn_k = int(d["neighbors"][0].shape[1])
        self.graph_knn: Optional[torch.cuda.CUDAGraph] = None
        self._knn_stream = None
        if tables == "device":
            # the tables are static buffers written by the builder's own graph on its own stream: in the pipelined engine the
            # (ALU-bound) table build of batch i+1 overlaps the (latency / HBM-bound) forward of batch i
            n = [p.shape[0] // self.B for p in self.inp["points"]]
            import ctypes
            self.knn_ws = ops._ws(_libmod.load().cofi_knn_pyramid_workspace((ctypes.c_int64 * len(n))(*n), len(n), self.B), dev)
            self._knn_stream = knn_stream if knn_stream is not None else torch.cuda.Stream(device=dev)
            self._tables_ready = torch.cuda.Event()
            self._tables_free = torch.cuda.Event()
            for key in ("neighbors", "subsampling", "upsampling"):   # never the batch's own tables: the builder fills them
                self.inp[key] = [torch.zeros_like(t) for t in self.inp[key]]
        self.img = batch["img"].to(dev).contiguous()
        self.kpt = torch.stack([k.to(torch.float32) for k in batch["fine_center_kpt_coors"]]).to(dev).contiguous()
        self.inline = torch.stack([k.to(torch.int64) for k in batch["fine_pc_inline_index"]]).to(dev).contiguous()
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.sim_stats = torch.zeros(2, dtype=torch.int32, device=dev)  # re-ranked candidates, full-scan rows (accumulated)
        self.out: Dict[str, torch.Tensor] = {}
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = 0
        self._host_in = None
        self._host_out = None
        self._stream = stream if stream is not None else torch.cuda.Stream(device=dev)
        self.use_graph = use_graph
        self._capture()

    def _capture(self):
        """warm-up (fills weight-pack / BN-fold / positional-encoding caches), then capture.  The graph bakes in parameter
        pointers and the packed / folded weight copies, so it is tied to the weights epoch (ops.weights_epoch) of this
        moment: run() re-captures when parameters were reloaded or updated by an optimiser step since."""
        dev = self.device
        self.graph = None
        self.epoch = ops.weights_epoch()
        self.graph_knn = None
        with torch.no_grad():
            with torch.cuda.stream(self._stream):
                for _ in range(2):
                    n0 = _libmod.launch_count()
                    self._knn_eager()
                    self._step_eager()
                    self.launches_per_step = _libmod.launch_count() - n0
                self._stream.synchronize()
                if self.use_graph:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self._stream):
                        self._step_eager()
                    self.graph = g
            if self.use_graph and self.tables == "device":
                with torch.cuda.stream(self._knn_stream):
                    gk = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gk, stream=self._knn_stream):
                        self._knn_eager()
                    self.graph_knn = gk
        torch.cuda.synchronize(dev)

    def _knn_eager(self):
        """tables="device": the 13 KNN tables of the batch from its point pyramid, into the static table buffers."""
        if self.tables != "device":
            return
        ops.knn_pyramid(self.inp["points"], frames=self.B, k=self.knn_k, k_up=1, workspace=self.knn_ws,
                        out={k: self.inp[k] for k in ("neighbors", "subsampling", "upsampling")})

    def run_knn(self, stream=None):
        """Enqueue the table build (tables="device") on `stream` (default: the engine's table stream)."""
        if self.tables != "device":
            return
        st = stream if stream is not None else self._knn_stream
        with torch.no_grad(), torch.cuda.stream(st):
            if self.graph_knn is not None:
                self.graph_knn.replay()
            else:
                self._knn_eager()

    # one forward over the static buffers; fills self.out
    def _step_eager(self):
        m, B = self.model, self.B
        core = m.core(self.inp, self.img, B)
        n1 = core["pc_decode_3"].shape[0] // B
        hw = m.pe_H * m.pe_W
        n4 = core["pc_norm"].shape[0] // B
        if self.mode == "test":
            t = m.tail_test_batched(core, self.inp["points"][-1], self.inp["points"][1], B, self.err, self.sim_stats)
            C = t["patch"].shape[2]
            fidx = ops.fine_match(t["patch"].view(B * n4, C, 16), t["fine_pc"].view(B * n4, C))   # eval_all.py:99-102
            self.out = {
                "img_norm": core["img_norm"], "pc_norm": core["pc_norm"],
                "img_score": core["img_score"], "pc_score": core["pc_score"],
                "patch": t["patch"], "fine_pc": t["fine_pc"], "fine_center_xy": t["fine_center_xy"],
                "coarse_pc_points": t["coarse_pc_points"], "count": t["count"], "sel": t["sel"],
                "fine_idx": fidx.view(B, n4),
            }
            return
        patches, fine = [], []
        for b in range(B):
            fine.append(ops.gather_rows(core["pc_decode_3"][b * n1:(b + 1) * n1], self.inline[b]))
            patches.append(ops.extract_patch(core["up2"], b, self.kpt[b], self.err))
        self.out = {
            # token layouts: [B*HW,128] / [B*N4,128]; public NCHW views are produced on demand (results())
            "img_norm": core["img_norm"], "pc_norm": core["pc_norm"],
            "img_score": core["img_score"], "pc_score": core["pc_score"],
            "patch": patches, "fine_pc": fine,
        }

    # ------------------------------------------------------------------------------------------ stepping
    def run(self):
        """One forward over whatever currently sits in the static input buffers (asynchronous)."""
        self._refresh()
        if self.tables == "device":  # stand-alone use: tables first (their own stream), then the forward
            self._tables_free.record(self._stream)
            self._knn_stream.wait_event(self._tables_free)      # the previous forward has finished reading the tables
            self.run_knn()
            self._tables_ready.record(self._knn_stream)
            self._stream.wait_event(self._tables_ready)
        self.run_forward()

    def _refresh(self):
        if self.epoch != ops.weights_epoch():
            self._capture()   # parameters changed since the capture: stale packs / folds / pointers must not be replayed

    def run_forward(self):
        """The forward over the static buffers (tables included), on the engine's compute stream."""
        self._refresh()
        with torch.no_grad():
            if self.graph is not None:
                with torch.cuda.stream(self._stream):
                    self.graph.replay()
            else:
                with torch.cuda.stream(self._stream):
                    self._step_eager()

    @property
    def stream(self):
        return self._stream

    def host_buffers(self, batch: Dict):
        """Pinned host copies of a batch (what a data loader would hand over)."""
        d = batch["pc_data_dict"]
        dev_tables = self.tables == "device"
        return {
            "points": [_pin(t) for t in d["points"]],
            "neighbors": [] if dev_tables else [_pin(t) for t in d["neighbors"]],
            "subsampling": [] if dev_tables else [_pin(t) for t in d["subsampling"]],
            "upsampling": [] if dev_tables else [_pin(t[:, :1]) for t in d["upsampling"]],
            "feats": _pin(d["feats"]), "img": _pin(batch["img"]),
            "kpt": _pin(torch.stack([k.to(torch.float32) for k in batch["fine_center_kpt_coors"]])),
            "inline": _pin(torch.stack([k.to(torch.int64) for k in batch["fine_pc_inline_index"]])),
        }

    def upload(self, host: Dict, stream=None) -> int:
        """Async H2D of one batch from pinned host buffers into the static inputs; returns the bytes copied."""
        nbytes = 0
        with torch.cuda.stream(stream if stream is not None else self._stream):
            for key in ("points", "neighbors", "subsampling", "upsampling"):
                for dst, src in zip(self.inp[key], host[key]):
                    dst.copy_(src, non_blocking=True)
                    nbytes += src.numel() * src.element_size()
            for dst, src in ((self.inp["feats"], host["feats"]), (self.img, host["img"]), (self.kpt, host["kpt"]),
                             (self.inline, host["inline"])):
                dst.copy_(src, non_blocking=True)
                nbytes += src.numel() * src.element_size()
        return nbytes

    def download(self, stream=None) -> int:
        """Async D2H of the step's results into pinned host buffers; returns the bytes copied."""
        if self.mode == "test":
            flat = [self.out[k] for k in ("img_norm", "pc_norm", "img_score", "pc_score", "patch", "fine_pc", "fine_center_xy",
                                          "coarse_pc_points", "count", "sel", "fine_idx")] + [self.err]
        else:
            flat = [self.out["img_norm"], self.out["pc_norm"], self.out["img_score"], self.out["pc_score"]] + \
                list(self.out["patch"]) + list(self.out["fine_pc"]) + [self.err]
        if self._host_out is None:
            self._host_out = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in flat]
        nbytes = 0
        with torch.cuda.stream(stream if stream is not None else self._stream):
            for dst, src in zip(self._host_out, flat):
                dst.copy_(src, non_blocking=True)
                nbytes += src.numel() * src.element_size()
        return nbytes

    def results(self) -> List[tuple]:
        """Reference-layout 8-tuples per frame (synchronises).  test mode: rows trimmed to each frame's match count."""
        self._stream.synchronize()
        m, B = self.model, self.B
        outs = []
        with torch.no_grad():
            if self.mode == "test":
                counts = self.out["count"][:, 0].tolist()
                if min(counts) < 4:
                    raise RuntimeError("fewer than 4 matches survive every threshold (the reference would loop forever)")
            for b in range(B):
                pub = m._public(self.out, b, B)
                if self.mode == "test":
                    n = counts[b]
                    outs.append(pub + (self.out["patch"][b, :n].clone(), self.out["fine_pc"][b, :n].clone(),
                                       self.out["fine_center_xy"][b, :, :n].clone(), self.out["coarse_pc_points"][b, :n].clone()))
                else:
                    outs.append(pub + (self.out["patch"][b], self.out["fine_pc"][b], None, None))
        assert int(self.err.item()) == 0, "extract_patch: window outside the feature map"
        return outs

    def correspondences(self):
        """test mode: per frame (imagePoints [n,2] fp32, objectPoints [n,3] fp32, fine index [n]) -- the inputs of the
        caller's PnP (evaluation/eval_all.py:99-107), assembled from the graph's fine-match output (synchronises)."""
        assert self.mode == "test"
        self._stream.synchronize()
        res = []
        counts = self.out["count"][:, 0].tolist()
        for b, n in enumerate(counts):
            idx = self.out["fine_idx"][b, :n]
            c = self.out["fine_center_xy"][b, :, :n]
            x = c[0] - 2 + torch.div(idx, 4, rounding_mode="floor")     # eval_all.py:103-105 (its x += idx//4 convention)
            y = c[1] - 2 + idx % 4
            res.append((torch.stack([x, y], 1), self.out["coarse_pc_points"][b, :n].clone(), idx.clone()))
        return res


    def correspondences_padded(self):
        """test mode, no host synchronisation: (imagePoints [B,N4,2] fp32, objectPoints [B,N4,3] fp32, match counts [B,2]
        int32) on the device -- the inputs of the batched pose step (cofii2p_b200.evaluate.solve_pose_batch); rows beyond a
        frame's count are padding."""
        assert self.mode == "test"
        with torch.no_grad(), torch.cuda.stream(self._stream):
            idx = self.out["fine_idx"].to(torch.float32)
            c = self.out["fine_center_xy"]
            x = c[:, 0] - 2 + torch.floor(idx / 4)       # eval_all.py:103-105 (its x += idx // 4 convention)
            y = c[:, 1] - 2 + idx % 4
            return torch.stack([x, y], 2).contiguous(), self.out["coarse_pc_points"], self.out["count"]


class PipelinedEngine:
    """Double-buffered end-to-end pipeline: while the graph of buffer set i computes, the H2D copy of batch i+1
    lands in buffer set (i+1)%2 on a copy stream and the D2H of batch i-1 drains on a third stream (PCIe is full
    duplex); with tables="device" the KNN-table graph of batch i+1 runs on a fourth stream beside the forward of batch i.
    Steady-state step time = max(H2D, compute, D2H) instead of their sum."""

    def __init__(self, model, batch: Dict, depth: int = 2, tables: str = "host", mode: str = "val"):
        dev = next(model.parameters()).device
        self.compute = torch.cuda.Stream(device=dev)
        self.h2d = torch.cuda.Stream(device=dev)
        self.d2h = torch.cuda.Stream(device=dev)
        self.knn = torch.cuda.Stream(device=dev)   # tables="device": table build of batch i+1 runs beside the forward of batch i
        self.tables = tables
        self.engines = [InferenceEngine(model, batch, mode=mode, use_graph=True, stream=self.compute, tables=tables,
                                        knn_stream=self.knn) for _ in range(depth)]
        self.tabled = [torch.cuda.Event() for _ in range(depth)]
        self.uploaded = [torch.cuda.Event() for _ in range(depth)]
        self.computed = [torch.cuda.Event() for _ in range(depth)]
        self.drained = [torch.cuda.Event() for _ in range(depth)]
        for e in self.computed + self.drained:
            e.record(self.compute)
        self.i = 0
        self.launches_per_step = self.engines[0].launches_per_step

    def step(self, host: Dict):
        """Enqueue one batch end to end (asynchronous); returns (h2d bytes, d2h bytes)."""
        k = self.i % len(self.engines)
        eng = self.engines[k]
        self.h2d.wait_event(self.computed[k])      # buffer set k is free once its previous compute finished
        nin = eng.upload(host, self.h2d)
        self.uploaded[k].record(self.h2d)
        if self.tables == "device":
            self.knn.wait_event(self.uploaded[k])  # (the upload itself waited for the previous forward over buffer set k)
            eng.run_knn(self.knn)
            self.tabled[k].record(self.knn)
            self.compute.wait_event(self.tabled[k])
        else:
            self.compute.wait_event(self.uploaded[k])
        self.compute.wait_event(self.drained[k])   # outputs of set k were copied out
        eng.run_forward()
        self.computed[k].record(self.compute)
        self.d2h.wait_event(self.computed[k])
        nout = eng.download(self.d2h)
        self.drained[k].record(self.d2h)
        self.i += 1
        return nin, nout

    def synchronize(self):
        self.h2d.synchronize()
        self.knn.synchronize()
        self.compute.synchronize()
        self.d2h.synchronize()

    def last_results(self):
        self.synchronize()
        return self.engines[(self.i - 1) % len(self.engines)].results()
