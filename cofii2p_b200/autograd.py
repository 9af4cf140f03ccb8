"""Differentiable versions of the hot-path ops: `torch.autograd.Function`s whose forward AND backward are the
hand-written kernels of libcofi_b200.so (ops.py).  Used by `cofii2p_b200.model` when the model is in training mode
with gradients enabled (reference train.py:189-285: model.train(); loss.backward()).  PyTorch contributes the
autograd graph bookkeeping and gradient accumulation only.

Contraction engine: whatever `ops.set_engine` selected (fp32 or tf32); the fp16 / fused-epilogue inference variants are
not used on the training path."""
from __future__ import annotations

import torch

from . import ops

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_SIGMOID = ops.ACT_NONE, ops.ACT_RELU, ops.ACT_LRELU, ops.ACT_SIGMOID


def active(module: torch.nn.Module) -> bool:
    """True when the differentiable path must be taken."""
    return module.training and torch.is_grad_enabled()


# ------------------------------------------------------------------------------------------------ linear
class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, act):
        y = ops.gemm(x, w, bias=b, act=act)
        ctx.act = act
        ctx.save_for_backward(x, w, y if act != ACT_NONE else None)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        g = dy.contiguous() if ctx.act == ACT_NONE else ops.act_bwd(dy, y, ctx.act)
        dx = dw = db = None
        n = w.shape[0]
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = ops.colsum(_pad_cols(g))[:n]
        if n % 4:                                          # the 1-channel score heads: contract over a zero-padded width
            g, w = _pad_cols(g), _pad_rows(w)
        if ctx.needs_input_grad[0]:
            dx = ops.gemm(g, ops.transpose2d(w))          # [M,N] x [K,N]^T
        if ctx.needs_input_grad[1]:
            dw = ops.gemm_tn(g, x)[:n]                     # dY^T X, no transposes
        return dx, dw, db, None


def _pad_cols(t):
    pad = -t.shape[1] % 4
    if not pad:
        return t
    out = torch.zeros((t.shape[0], t.shape[1] + pad), dtype=t.dtype, device=t.device)
    out[:, :t.shape[1]] = t
    return out


def _pad_rows(t):
    pad = -t.shape[0] % 4
    if not pad:
        return t
    out = torch.zeros((t.shape[0] + pad, t.shape[1]), dtype=t.dtype, device=t.device)
    out[:t.shape[0]] = t
    return out


def linear(x, w, b=None, act: int = ACT_NONE):
    return _Linear.apply(x.contiguous(), w, b, act)


class _LinearStats(torch.autograd.Function):
    """Linear whose GEMM epilogue also emits the per-128-row-tile column statistics of the output (tf32 engine): the
    GroupNorm that follows (norm_rows_pre) needs no statistics pass over the activation.  Backward = _Linear's."""

    @staticmethod
    def forward(ctx, x, w, b):
        y, st = ops.gemm_colstats(x, w, bias=b)
        ctx.act = ACT_NONE
        ctx.save_for_backward(x, w, None)
        ctx.has_bias = b is not None
        ctx.mark_non_differentiable(st)
        return y, st

    @staticmethod
    def backward(ctx, dy, _dst):
        return _Linear.backward(ctx, dy)[:3]


def linear_stats(x, w, b=None):
    return _LinearStats.apply(x.contiguous(), w, b)


# ------------------------------------------------------------------------------------------------ KPConv
class _KPConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, weights, bias, q_points, s_points, nbr, kernel_points, sigma, frames, kp_reach, want_stats=False):
        K, C, Co = weights.shape
        packed = ops.pack_points(s_points, feats)
        agg, cnt = ops.kpconv_aggregate(feats, packed, q_points, nbr, kernel_points, sigma, frames, kp_reach)
        wt = ops.transpose2d(weights.reshape(K * C, Co))
        # the aggregate [M, K*C] and the packed points are kept for the backward instead of being recomputed: a few GB
        # per step, nothing next to 180 GB of HBM, and no extra traffic (the tensors exist anyway)
        ctx.save_for_backward(feats, weights, q_points, s_points, nbr, kernel_points, cnt, agg, packed)
        ctx.meta = (float(sigma), int(frames), float(kp_reach), bias is not None)
        if want_stats:  # the weight-apply GEMM's epilogue feeds the GroupNorm that follows
            y, st = ops.gemm_colstats(agg, wt, bias=bias, rowdiv=cnt)
            ctx.mark_non_differentiable(st)
            return y, st
        return ops.gemm(agg, wt, bias=bias, rowdiv=cnt)

    @staticmethod
    def backward(ctx, dy, _dst=None):
        feats, weights, q_points, s_points, nbr, kernel_points, cnt, agg, packed = ctx.saved_tensors
        sigma, frames, kp_reach, has_bias = ctx.meta
        K, C, Co = weights.shape
        dy = dy.contiguous()
        g = ops.rowscale(dy, cnt)                                   # d(acc) = dY / cnt
        dfeats = dw = db = None
        if ctx.needs_input_grad[1]:
            dw = ops.gemm_tn(agg, g).view(K, C, Co)                 # agg^T g = d weights.reshape(K*C, Co)
        if ctx.needs_input_grad[0]:
            dagg = ops.gemm(g, weights.reshape(K * C, Co))          # g W^T with W stored as [K*C, Co]
            dfeats = ops.kpconv_aggregate_bwd(dagg, C, packed, q_points, nbr, kernel_points, sigma, frames, kp_reach)
        if has_bias and ctx.needs_input_grad[2]:
            db = ops.colsum(dy)
        return dfeats, dw, db, None, None, None, None, None, None, None, None


def kpconv(feats, weights, bias, q_points, s_points, nbr, kernel_points, sigma, frames, kp_reach, want_stats=False):
    """want_stats (tf32 engine, whole 128-row tiles per frame): returns (y, tile statistics) for norm_rows_pre."""
    return _KPConv.apply(feats.contiguous(), weights, bias, q_points, s_points, nbr, kernel_points, sigma, frames, kp_reach,
                         bool(want_stats))


# ------------------------------------------------------------------------------------------------ normalisations
class _NormRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, residual, frames, groups, eps, act):
        y, mean, var = ops.norm_rows(x, frames, groups, gamma, beta, eps, residual=residual, act=act, want_stats=True)
        mr = torch.stack((mean.reshape(-1), torch.rsqrt(var.reshape(-1) + eps)), 1)   # (mean, rstd) per (frame, group)
        ctx.save_for_backward(x, gamma, y, mr)
        ctx.meta = (frames, groups, eps, act, residual is not None)
        ctx.mark_non_differentiable(mean, var)
        return y, mean, var

    @staticmethod
    def backward(ctx, dy, _dmean, _dvar):
        x, gamma, y, mr = ctx.saved_tensors
        frames, groups, eps, act, has_res = ctx.meta
        dx, dres, dgamma, dbeta = ops.norm_rows_bwd(x, dy.contiguous(), y, mr, frames, groups, gamma, act, has_res)
        return dx, dgamma, dbeta, dres, None, None, None, None


class _NormRowsPre(torch.autograd.Function):
    """GroupNorm on statistics that came out of the producing GEMM's epilogue; backward identical to _NormRows."""

    @staticmethod
    def forward(ctx, x, tile_stats, gamma, beta, residual, frames, groups, eps, act):
        y, mr = ops.norm_rows_pre(x, tile_stats, frames, groups, gamma, beta, eps, residual=residual, act=act, return_mr=True)
        ctx.save_for_backward(x, gamma, y, mr)
        ctx.meta = (frames, groups, eps, act, residual is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, y, mr = ctx.saved_tensors
        frames, groups, eps, act, has_res = ctx.meta
        dx, dres, dgamma, dbeta = ops.norm_rows_bwd(x, dy.contiguous(), y, mr, frames, groups, gamma, act, has_res)
        return dx, None, dgamma, dbeta, dres, None, None, None, None


def norm_rows_pre(x, tile_stats, frames, groups, gamma, beta, eps=1e-5, residual=None, act=ACT_NONE):
    return _NormRowsPre.apply(x.contiguous(), tile_stats, gamma, beta, None if residual is None else residual.contiguous(),
                              frames, groups, eps, act)


def norm_rows(x, frames, groups, gamma=None, beta=None, eps=1e-5, residual=None, act=ACT_NONE, return_stats=False):
    """return_stats: also the per-(frame, group) mean and biased variance the forward computed (BatchNorm running stats)."""
    y, mean, var = _NormRows.apply(x.contiguous(), gamma, beta, None if residual is None else residual.contiguous(), frames,
                                   groups, eps, act)
    return (y, mean, var) if return_stats else y


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, residual, eps, act):
        y = ops.layer_norm_rows(x, gamma, beta, eps, act=act, residual=residual)
        ctx.save_for_backward(x, gamma, beta)
        ctx.meta = (eps, act, residual is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta = ctx.saved_tensors
        eps, act, has_res = ctx.meta
        dx, dgamma, dbeta = ops.layer_norm_bwd(x, dy, gamma, beta, eps, act)
        return dx, dgamma, dbeta, (dy if has_res else None), None, None


def layer_norm(x, gamma, beta, eps=1e-5, act=ACT_NONE, residual=None):
    return _LayerNorm.apply(x.contiguous(), gamma, beta, None if residual is None else residual.contiguous(), eps, act)


class _L2Norm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, add):
        ctx.save_for_backward(x)
        ctx.has_add = add is not None
        return ops.l2norm_rows(x, add=add)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.l2norm_bwd(x, dy), (dy if ctx.has_add else None)


def l2norm(x, add=None):
    return _L2Norm.apply(x.contiguous(), add)


class _ColNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, frames):
        ctx.save_for_backward(x)
        ctx.frames = frames
        return ops.colnorm_rows(x, frames)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.colnorm_bwd(x, dy, ctx.frames), None


def colnorm(x, frames):
    return _ColNorm.apply(x.contiguous(), frames)


# ------------------------------------------------------------------------------------------------ gathers
class _Gather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx, idx_stride, frames, rows_out):
        ctx.save_for_backward(idx)
        ctx.meta = (idx_stride, frames, x.shape[0])
        return ops.gather_rows(x, idx, idx_stride=idx_stride, frames=frames, rows_out=rows_out)

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        idx_stride, frames, rows_src = ctx.meta
        return ops.scatter_add_rows(dy, idx, idx_stride, frames, rows_src), None, None, None, None


def gather(x, idx, idx_stride=1, frames=1, rows_out=None):
    return _Gather.apply(x.contiguous(), idx, idx_stride, frames, rows_out)


class _MaxpoolRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, nbr, frames):
        ctx.save_for_backward(x, nbr)
        ctx.frames = frames
        return ops.maxpool_rows(x, nbr, frames)

    @staticmethod
    def backward(ctx, dy):
        x, nbr = ctx.saved_tensors
        return ops.maxpool_rows_bwd(x, nbr, dy, ctx.frames), None, None


def maxpool_rows(x, nbr, frames):
    return _MaxpoolRows.apply(x.contiguous(), nbr, frames)


# ------------------------------------------------------------------------------------------------ attention
class _Attention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, frames, heads, scale):
        out, lse = ops.attention_fwd_lse(q, k, v, frames, heads, scale)
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.meta = (frames, heads, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse = ctx.saved_tensors
        frames, heads, scale = ctx.meta
        dq, dk, dv = ops.attention_bwd(q, k, v, out, dout, lse, frames, heads, scale)
        return dq, dk, dv, None, None, None


def attention(q, k, v, frames, heads, scale):
    return _Attention.apply(q.contiguous(), k.contiguous(), v.contiguous(), frames, heads, scale)


# ------------------------------------------------------------------------------------------------ image stream
def _pack_w(w, cin_pad):
    """[Co,Ci,kh,kw] -> [Co, kh*kw*Cip] (ci fastest), zero-padded input channels."""
    wp = w.permute(0, 2, 3, 1)
    if cin_pad > wp.shape[3]:
        wp = torch.nn.functional.pad(wp, (0, cin_pad - wp.shape[3]))
    return wp.reshape(w.shape[0], -1).contiguous()


class _Conv2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, stride, pad):
        Co, Ci, kh, kw = w.shape
        cip = x.shape[3]
        y = ops.conv2d_nhwc(x, _pack_w(w.detach(), cip), kh, kw, stride, pad)
        ctx.save_for_backward(x, w)
        ctx.meta = (stride, pad)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        stride, pad = ctx.meta
        Co, Ci, kh, kw = w.shape
        cip = x.shape[3]
        dy = dy.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[1]:
            dwp = ops.conv2d_wgrad_nhwc(x, dy, kh, kw, stride, pad)                       # [Co, kh*kw*Cip]
            dw = dwp.view(Co, kh, kw, cip)[..., :Ci].permute(0, 3, 1, 2).contiguous()
        if ctx.needs_input_grad[0]:
            # input gradient = stride-1 correlation of the (zero-dilated) output gradient with the flipped kernel
            g = dy if stride == 1 else ops.dilate2_nhwc(dy)
            wt = w.detach().flip(2, 3).permute(1, 2, 3, 0)                                # [Ci, kh, kw, Co]
            if cip > Ci:
                wt = torch.nn.functional.pad(wt, (0, 0, 0, 0, 0, 0, 0, cip - Ci))
            dx = ops.conv2d_nhwc(g, wt.reshape(cip, -1).contiguous(), kh, kw, 1, kh - 1 - pad)
            if tuple(dx.shape) != tuple(x.shape):                                        # odd sizes: crop / pad to the input
                dx = dx[:, :x.shape[1], :x.shape[2], :].contiguous()
        return dx, dw, None, None


def conv2d(x_nhwc, w, stride, pad):
    return _Conv2d.apply(x_nhwc.contiguous(), w, stride, pad)


class _Maxpool2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.maxpool2d_3x3s2_nhwc(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.maxpool2d_3x3s2_bwd(x, dy)


def maxpool2d(x):
    return _Maxpool2d.apply(x.contiguous())


class _UpsampleCat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x1, x2):
        ctx.c1 = x1.shape[3]
        return ops.upsample2x_cat_nhwc(x1, x2)

    @staticmethod
    def backward(ctx, dy):
        return ops.upsample2x_cat_bwd(dy, ctx.c1)


def upsample2x_cat(x1, x2):
    return _UpsampleCat.apply(x1.contiguous(), x2.contiguous())


class _ToNCHW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return ops.nhwc_to_nchw(x)

    @staticmethod
    def backward(ctx, dy):
        return ops.nchw_to_nhwc(dy.contiguous())


def nhwc_to_nchw(x):
    return _ToNCHW.apply(x.contiguous())


class _ExtractPatch(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fmap, b, centers, err):
        ctx.save_for_backward(centers)
        ctx.meta = (tuple(fmap.shape), b)
        return ops.extract_patch(fmap, b, centers, err)

    @staticmethod
    def backward(ctx, dpatch):
        (centers,) = ctx.saved_tensors
        shape, b = ctx.meta
        return ops.extract_patch_bwd(dpatch, shape, b, centers), None, None, None


def extract_patch(fmap, b, centers, err):
    return _ExtractPatch.apply(fmap.contiguous(), b, centers.contiguous(), err)


class _ExtractPatchBatched(torch.autograd.Function):
    """All frames in one launch each way: fmap [B,H,W,C], centers [B,2,n] -> [B,n,C,4,4]; one zero-initialised map gradient
    for the whole batch (the per-frame form returns a full-size map gradient per frame for autograd to sum)."""

    @staticmethod
    def forward(ctx, fmap, centers, err):
        ctx.save_for_backward(centers)
        ctx.shape = tuple(fmap.shape)
        return ops.extract_patch_batched(fmap, centers, err)

    @staticmethod
    def backward(ctx, dpatch):
        (centers,) = ctx.saved_tensors
        return ops.extract_patch_batched_bwd(dpatch, ctx.shape, centers), None, None


def extract_patch_batched(fmap, centers, err):
    return _ExtractPatchBatched.apply(fmap.contiguous(), centers.contiguous(), err)


# ------------------------------------------------------------------------------------------ fused training losses
class _DescLoss(torch.autograd.Function):
    """sum over frames of desc_loss (reference model/loss.py:69-93) from token-layout descriptor maps + row indices."""

    @staticmethod
    def forward(ctx, img_tok, pc_tok, pix, kpt, mask, frames, pos_margin, neg_margin, log_scale):
        loss, _, d_img, d_pc = ops.desc_loss_fwd(img_tok, pix, pc_tok, kpt, mask, frames, pos_margin, neg_margin, log_scale)
        ctx.save_for_backward(d_img, d_pc, pix, kpt)
        ctx.meta = (frames, img_tok.shape[0] // frames, pc_tok.shape[0] // frames)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        d_img, d_pc, pix, kpt = ctx.saved_tensors
        frames, img_rows, pc_rows = ctx.meta
        g = dloss.contiguous()   # [frames]: the upstream gradient of every frame's loss
        return (ops.scatter_scaled_rows(d_img, pix, img_rows, frames, g), ops.scatter_scaled_rows(d_pc, kpt, pc_rows, frames, g),
                None, None, None, None, None, None, None)


def desc_loss_tokens(img_tok, pc_tok, pix, kpt, mask, frames, pos_margin, neg_margin, log_scale=10.0):
    return _DescLoss.apply(img_tok, pc_tok, pix, kpt, mask, frames, pos_margin, neg_margin, log_scale)


class _OverlapLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, score_tok, idx, n_in, frames):
        loss, d = ops.overlap_loss_fwd(score_tok, idx, n_in, frames)
        ctx.save_for_backward(d, idx)
        ctx.meta = (frames, score_tok.numel() // frames, tuple(score_tok.shape))
        return loss

    @staticmethod
    def backward(ctx, dloss):
        d, idx = ctx.saved_tensors
        frames, rows, shape = ctx.meta
        g = dloss.contiguous()
        return ops.scatter_scaled_rows(d.reshape(-1, 1), idx, rows, frames, g).view(shape), None, None, None


def overlap_loss_tokens(score_tok, idx, n_in, frames):
    return _OverlapLoss.apply(score_tok, idx, n_in, frames)


class _FineCircleLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, patch, fpc, rel, frames, bad_flag):
        loss, d_patch, d_fpc = ops.fine_circle_loss_fwd(patch, fpc, rel, frames, bad_flag)
        ctx.save_for_backward(d_patch, d_fpc)
        ctx.frames = frames
        return loss

    @staticmethod
    def backward(ctx, dloss):
        d_patch, d_fpc = ctx.saved_tensors
        g = dloss.contiguous()
        return (ops.scatter_scaled_rows(d_patch, None, 0, ctx.frames, g), ops.scatter_scaled_rows(d_fpc, None, 0, ctx.frames, g),
                None, None, None)


def fine_circle_loss_rows(patch, fpc, rel, frames, bad_flag=None):
    return _FineCircleLoss.apply(patch, fpc, rel, frames, bad_flag)
