"""Data-parallel training step of the hot path (BASELINE.json configs[4]).

What the reference does per iteration (reference train.py:186-285): one frame through `CoFiI2P.forward(mode='train')`,
three losses (`model/loss.py`), `loss.backward()`, `torch.optim.Adam.step()` -- single GPU, batch 1.  Here:

* `training_losses` restates train.py:233-283 (the gathers that pick the supervised rows and the three losses);
* `TrainStep` runs B stacked frames per rank through `CoFiI2P.forward_batch(mode='train')`.  Every normalisation keeps
  per-frame statistics, so the step equals B reference iterations whose gradients are averaged (gradient accumulation)
  -- the only batching the reference's batch-1 model admits;
* forward and backward are the library's kernels (cofii2p_b200/autograd.py); parameters, gradients and the Adam moments
  live in three flat fp32 buffers: one NCCL all-reduce over NVLink for the gradient, one fused Adam kernel for the update;
* frames shard across ranks (cofii2p_b200/shard.py); there is no other collective on the path.
"""
from typing import Dict, List, Optional

import torch

from . import autograd as ad
from . import ops
from .model.loss import desc_loss_algebra as desc_loss, fine_circle_loss_algebra as fine_circle_loss, \
    overlap_loss_algebra as overlap_loss

__all__ = ["training_losses", "fused_training_losses", "TrainStep"]


def fused_training_losses(model, batch: Dict, opt):
    """The same three losses for ALL B stacked frames of a step through the fused forward + backward kernels (csrc/loss.cu),
    straight from the token-layout network outputs: no per-frame NCHW transposes, fancy-index gathers or autograd graphs of
    ~60 element-wise ops per frame.  Returns (mean loss over frames, [desc, overlap, fine] means, bad-supervision flag).
    What stays tensor algebra is the 64 x 64 projection mask of train.py:237-251 (a dozen tiny batched launches per step)."""
    B = batch["frames"]
    core, patch, fine_pc = model.forward_train_tokens(batch)
    dev = core["pc_norm"].device
    n4 = core["pc_norm"].shape[0] // B
    W = model.pe_W
    kpt = torch.stack(batch["pc_kpt_idx"]).to(dev)                                       # [B, n]
    outl = torch.stack(batch["pc_outline_idx"]).to(dev)
    pix = torch.stack(batch["coarse_img_kpt_idx"]).to(dev)
    P, K4 = torch.stack(batch["P"]).to(dev), torch.stack(batch["K_4"]).to(dev)           # [B,4,4], [B,3,3]
    pts4 = batch["pc_data_dict"]["points"][-1].view(B, n4, 3)
    xyz = torch.gather(pts4, 1, kpt.unsqueeze(-1).expand(-1, -1, 3)).transpose(1, 2)     # [B,3,n]            train.py:237
    img_xy = torch.stack(((pix % W).to(torch.float32), torch.div(pix, W, rounding_mode="floor").to(torch.float32)), 1)  # :245
    proj = K4 @ (P[:, 0:3, 0:3] @ xyz + P[:, 0:3, 3:])                                   # :247
    proj_xy = proj[:, 0:2] / proj[:, 2:]
    mask = (torch.sqrt(torch.sum(torch.square(img_xy.unsqueeze(-1) - proj_xy.unsqueeze(-2)), dim=1))
            <= opt.dist_thres).float()                                                   # [B, n_img, n_pc]   :251
    l_desc = ad.desc_loss_tokens(core["img_norm"], core["pc_norm"], pix, kpt, mask, B, opt.pos_margin, opt.neg_margin)   # :254
    l_coarse = ad.overlap_loss_tokens(core["pc_score"], torch.cat((kpt, outl), 1), kpt.shape[1], B)                      # :256-260
    fine_xy = torch.stack(batch["fine_xy"]).to(dev)
    centers = torch.stack([k.to(torch.float32) for k in batch["fine_center_kpt_coors"]]).to(dev)
    rel = torch.floor(fine_xy) - centers + 2                                             # [B,2,n]            :267
    bad = ((rel < 0) | (rel > 3)).any().to(torch.int32)
    rel_index = (rel[:, 1] * 4 + rel[:, 0]).long().clamp_(0, 15).reshape(-1)
    l_fine = ad.fine_circle_loss_rows(patch, fine_pc, rel_index, B)                      # :283
    parts = torch.stack((l_desc.detach().mean(), l_coarse.detach().mean(), l_fine.detach().mean()))
    return (l_desc.sum() + l_coarse.sum() + l_fine.sum()) / B, parts, bad


def training_losses(out, sup: Dict, opt, points4: torch.Tensor):
    """Losses of one frame (reference train.py:233-283) as plain tensor algebra on the model's public outputs -- the
    definition the fused path (fused_training_losses) is tested against. `out` is the model's 8-tuple in train mode, `sup` the frame's
    supervision (pc_kpt_idx, pc_outline_idx, coarse_img_kpt_idx, K_4, P, fine_xy, fine_center_kpt_coors), `points4`
    the frame's coarsest-level points [n4,3]."""
    img_features, pc_features, _, coarse_pc_score, fine_patch, fine_pc = out[:6]
    dev = pc_features.device
    n = opt.num_kpt
    kpt, outl, pix = sup["pc_kpt_idx"], sup["pc_outline_idx"], sup["coarse_img_kpt_idx"]
    pc_in = pc_features[:, kpt]                                                          # :233
    xyz_in = points4[kpt].t()                                                            # :237
    C, H, W = img_features.shape[1:]
    img_in = img_features.reshape(C, H * W)[:, pix]                                      # :239-243
    gx = (pix % W).to(torch.float32)
    gy = torch.div(pix, W, rounding_mode="floor").to(torch.float32)
    img_xy = torch.stack((gx, gy), 0)                                                    # :245
    P, K4 = sup["P"], sup["K_4"]
    proj = K4 @ (P[0:3, 0:3] @ xyz_in + P[0:3, 3:])                                      # :247
    proj_xy = proj[0:2] / proj[2:]
    mask = (torch.sqrt(torch.sum(torch.square(img_xy.unsqueeze(-1) - proj_xy.unsqueeze(-2)), dim=0))
            <= opt.dist_thres).float()                                                   # :251
    loss_desc, _ = desc_loss(dev, img_in, pc_in, mask, pos_margin=opt.pos_margin, neg_margin=opt.neg_margin)   # :254
    score = coarse_pc_score.reshape(-1)
    loss_coarse = overlap_loss(dev, score[kpt], score[outl])                             # :256-260
    rel = torch.floor(sup["fine_xy"]) - sup["fine_center_kpt_coors"].to(torch.float32) + 2   # :267 (integer pixels)
    # the reference indexes label[arange, rel_index] directly and raises when the supervised pixel falls outside the 4x4
    # window (:267-283).  Inside a captured graph nothing can raise, so the violation is recorded in a device flag that
    # TrainStep.check_errors() surfaces; the index is clamped only to keep the scatter in bounds.
    bad = ((rel < 0) | (rel > 3)).any().to(torch.int32)
    rel_index = (rel[1] * 4 + rel[0]).long().clamp_(0, 15)
    loss_fine = fine_circle_loss(dev, fine_patch, fine_pc, rel_index, n)                 # :283
    return loss_desc + loss_coarse + loss_fine, (loss_desc.detach(), loss_coarse.detach(), loss_fine.detach()), bad


class TrainStep:
    """One optimisation step over B stacked frames per rank (Adam, reference train.py:156 defaults)."""

    def __init__(self, model, opt, lr: Optional[float] = None, betas=(0.9, 0.999), eps: float = 1e-8, group=None,
                 fused_losses: bool = True):
        """fused_losses: the three losses of all frames through csrc/loss.cu (fused_training_losses); False = the reference's
        formulas as tensor algebra, frame by frame (training_losses) -- the path the fused kernels are tested against."""
        self.model, self.opt = model, opt
        self.fused_losses = fused_losses
        self.lr = float(opt.lr if lr is None else lr)
        self.betas, self.eps = betas, eps
        self.group = group
        self.world = torch.distributed.get_world_size(group) if torch.distributed.is_initialized() else 1
        self.step_count = 0
        self.live: Optional[List[torch.nn.Parameter]] = None
        self.flat_p = self.flat_g = self.m = self.v = None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self._static = self._g_loss = self._g_parts = self._g_err = self._g_bad = None

    # -------------------------------------------------------------------------------------------------------------
    def _flatten(self, live: List[torch.nn.Parameter]) -> None:
        """Move the live parameters into one flat buffer (each parameter becomes a view of it), and give every one a
        view of the flat gradient buffer as its `.grad`, so autograd accumulates in place and the all-reduce and
        Adam each touch one contiguous range.  Parameters that never receive a gradient (about 40 % of the
        reference's 51.6 M: dead layers) stay where they are, as under torch.optim.Adam."""
        dev = live[0].device
        sizes = [(p.numel() + 3) // 4 * 4 for p in live]                                 # 16-byte aligned slots
        total = sum(sizes)
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.m, self.v = torch.zeros_like(self.flat_p), torch.zeros_like(self.flat_p)
        off = 0
        for p, sz in zip(live, sizes):
            n = p.numel()
            self.flat_p[off:off + n].copy_(p.data.reshape(-1))
            g = self.flat_g[off:off + n].view_as(p)
            if p.grad is not None:
                g.copy_(p.grad)
            p.data = self.flat_p[off:off + n].view_as(p)
            p.grad = g
            off += sz
        self.live = live
        # parameter storage moved: packed-weight caches and captured inference graphs hold the old pointers
        ops.bump_weights_epoch()
        if hasattr(self.model, "_graphs"):
            self.model._graphs = {}
        self._sync_replicas()

    def _sync_replicas(self) -> None:
        """Data parallel: every rank must hold the same flat layout and start from the same values.  The layout (names and
        sizes of the live parameters, in order) is hashed and compared across ranks -- if the live sets differed, the
        all-reduce would mix unrelated slots -- then parameters and all buffers (BatchNorm running statistics) are
        broadcast from rank 0."""
        if self.world <= 1:
            return
        import hashlib
        dist = torch.distributed
        names = {id(p): n for n, p in self.model.named_parameters()}
        desc = ";".join(f"{names.get(id(p), '?')}:{p.numel()}" for p in self.live)
        h = int.from_bytes(hashlib.sha256(desc.encode()).digest()[:7], "little")
        mine = torch.tensor([h, self.flat_p.numel()], dtype=torch.int64, device=self.flat_p.device)
        lo, hi = mine.clone(), mine.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
        if not (torch.equal(lo, mine) and torch.equal(hi, mine)):
            raise RuntimeError("TrainStep: the live-parameter layout differs between ranks (different parameters received a "
                               "gradient on the first step); the flat gradient all-reduce would mix unrelated slots")
        dist.broadcast(self.flat_p, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        for b in self.model.buffers():
            dist.broadcast(b, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)

    def loss(self, batch: Dict, check: bool = True):
        """Mean over the rank's frames of the reference's per-frame loss."""
        model, B = self.model, batch["frames"]
        if self.fused_losses:
            total, parts, self.last_bad_supervision = fused_training_losses(model, batch, self.opt)
            if check and int(model.last_err.item()) != 0:
                raise AssertionError("extract_patch: a 4x4 window falls outside the feature map")
            return total, parts
        outs = model.forward_batch(batch, "train", check=check)
        n4 = batch["pc_data_dict"]["points"][-1].shape[0] // B
        total, parts = None, []
        for b in range(B):
            sup = {k: batch[k][b] for k in ("pc_kpt_idx", "pc_outline_idx", "coarse_img_kpt_idx", "K_4", "P", "fine_xy",
                                           "fine_center_kpt_coors")}
            l, p, bad = training_losses(outs[b], sup, self.opt, batch["pc_data_dict"]["points"][-1][b * n4:(b + 1) * n4])
            total = l if total is None else total + l
            parts.append(torch.stack(p))
            flags = bad if b == 0 else flags + bad
        self.last_bad_supervision = flags
        return total / B, torch.stack(parts).mean(0)

    def backward(self, batch: Dict):
        """forward + loss + backward; the rank's gradient is left in the parameters' `.grad` (flat after step 1)."""
        self.model.train()
        if self.flat_g is not None:
            self.flat_g.zero_()
        else:
            for p in self.model.parameters():
                p.grad = None
        loss, parts = self.loss(batch)
        loss.backward()
        if self.live is None:
            self._flatten([p for p in self.model.parameters() if p.requires_grad and p.grad is not None])
            if self.world > 1:  # that gradient predates the rank-0 broadcast of _sync_replicas: recompute it once
                self.flat_g.zero_()
                loss, parts = self.loss(batch)
                loss.backward()
        return loss.detach(), parts

    # -------------------------------------------------------------------------------------------------------------
    def enable_cuda_graph(self, batch: Dict) -> None:
        """Capture forward + losses + backward over `batch`'s device buffers into one CUDA graph (about 2,000 kernel
        launches and the autograd bookkeeping between them leave the critical path).  Later steps copy their batch into
        these buffers and replay.  The gradient all-reduce and the Adam kernel (whose bias correction takes the step
        number as an argument) stay outside the graph.  Needs the flat buffers, i.e. at least one eager step."""
        if self.live is None:
            self.backward(batch)
        self.model.train()
        # the eager step created the AccumulateGrad nodes on the default stream; the capture runs on its own stream
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # warm-up on the capture-side stream (allocator, caches)
            self.flat_g.zero_()
            l, _ = self.loss(batch, check=False)
            l.backward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.flat_g.zero_()
            l, parts = self.loss(batch, check=False)
            l.backward()
            self._g_loss, self._g_parts, self._g_err = l.detach(), parts, self.model.last_err
            self._g_bad = self.last_bad_supervision
        self.graph, self._static = g, batch

    def _load_static(self, batch: Dict) -> None:
        def cp(dst, src):
            if torch.is_tensor(dst):
                dst.copy_(src, non_blocking=True)
            elif isinstance(dst, list):
                for d, s in zip(dst, src):
                    cp(d, s)
            elif isinstance(dst, dict):
                for k in dst:
                    cp(dst[k], src[k])
        cp(self._static, batch)

    def check_errors(self) -> None:
        """Host-side check of the captured step's out-of-map flag (synchronises)."""
        if self._g_err is not None and int(self._g_err.item()) != 0:
            raise AssertionError("extract_patch: a 4x4 window falls outside the feature map")
        bad = self._g_bad if self.graph is not None else getattr(self, "last_bad_supervision", None)
        if bad is not None and int(bad.item()) != 0:
            raise IndexError("fine supervision outside the 4x4 patch window (the reference's label[...] indexing raises, "
                             "train.py:267-283)")

    def step(self, batch: Dict):
        if self.graph is not None:
            if batch is not self._static:
                self._load_static(batch)
            self.graph.replay()
            loss, parts = self._g_loss, self._g_parts
        else:
            loss, parts = self.backward(batch)
        if self.world > 1:
            torch.distributed.all_reduce(self.flat_g, group=self.group)                  # NCCL sum over NVLink
        self.step_count += 1
        ops.adam_step(self.flat_p, self.flat_g, self.m, self.v, self.lr, self.betas[0], self.betas[1], self.eps,
                      self.step_count, grad_scale=1.0 / self.world)
        # the raw-pointer update (and the replayed BatchNorm running-stat updates) are invisible to tensor._version:
        # invalidate every cache derived from parameter values (packed weights, BN folds, captured inference graphs)
        ops.bump_weights_epoch()
        return loss, parts
