"""LoFTR-style I2P transformer (reference model/transformer/transformer.py:15-103) on the B200 kernels.
Tokens are row-major [B*L, C]; batch handling is by frame segments."""
import copy

import torch
import torch.nn as nn

from ... import autograd as ad
from ... import ops
from .linear_attention import FullAttention


def _attn_tc(rows: int) -> bool:
    """tcgen05 flash attention for the "tr_attn" precision group (V is then produced directly as V^T)."""
    with ops.group("tr_attn"):
        return ops.engine_id() == ops.ENGINE_TF32 and rows % 4 == 0


class LoFTREncoderLayer(nn.Module):
    def __init__(self, d_model, nhead, attention="full"):
        super().__init__()
        assert attention == "full", "the reference instantiates ATTENTION='full' (model/network.py:35)"
        self.dim, self.nhead = d_model // nhead, nhead
        for name in ("q_proj", "k_proj", "v_proj"):  # bias-free projections (state_dict order q, k, v, merge, mlp, norms)
            setattr(self, name, nn.Linear(d_model, d_model, bias=False))
        self.attention = FullAttention()
        self.merge = nn.Linear(d_model, d_model, bias=False)
        hidden = 2 * d_model
        self.mlp = nn.Sequential(nn.Linear(hidden, hidden, bias=False), nn.ReLU(True), nn.Linear(hidden, d_model, bias=False))
        self.norm1, self.norm2 = nn.LayerNorm(d_model), nn.LayerNorm(d_model)

    def forward(self, x, source, frames: int = 1):
        """x [B*L,C], source [B*S,C] (None = x itself) -> [B*L,C]."""
        C = x.shape[1]
        if source is None:
            source = x
        if ad.active(self):  # training: differentiable kernels, attention forward with saved log-sum-exp
            q = ad.colnorm(ad.linear(x, self.q_proj.weight), frames)
            k = ad.linear(source, self.k_proj.weight)
            v = ad.linear(source, self.v_proj.weight)
            msg = ad.attention(q, k, v, frames, self.nhead, 1.0 / self.dim ** 0.5)
            m = ad.layer_norm(ad.linear(msg, self.merge.weight), self.norm1.weight, self.norm1.bias, self.norm1.eps)
            h = ad.linear(torch.cat([x, m], 1), self.mlp[0].weight, None, ops.ACT_RELU)
            h = ad.linear(h, self.mlp[2].weight)
            return ad.layer_norm(h, self.norm2.weight, self.norm2.bias, self.norm2.eps, residual=x)
        # F.normalize(q) with default dim=1 == L2 over the sequence axis per (head, channel) (reference :53)
        q = ops.colnorm_rows(ops.gemm(x, self.q_proj.weight), frames)
        k = ops.gemm(source, self.k_proj.weight)
        if _attn_tc(source.shape[0]):
            # tcgen05 flash attention: V is produced directly as V^T (K-major operand) by swapping GEMM operands
            vt = ops.gemm(self.v_proj.weight, source)
            msg = ops.attention_vt(q, k, vt, frames, self.nhead, 1.0 / self.dim ** 0.5)
        else:
            v = ops.gemm(source, self.v_proj.weight)
            with ops.group("tr_attn"):
                msg = self.attention(q, k, v, frames, self.nhead)
        # message = norm1(merge(message)): Linear + LayerNorm fused in the GEMM epilogue
        m = ops.gemm_ln(msg, self.merge.weight, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        # mlp[0] on cat([x, message]) without materialising the concat: two K=C GEMMs accumulating into one output
        w1 = self.mlp[0].weight
        h = ops.gemm(x, w1[:, :C])
        h = ops.gemm(m, w1[:, C:], out=h, accumulate=True, act=ops.ACT_RELU)
        # x + norm2(mlp[2](h)): Linear + LayerNorm + residual in one kernel
        return ops.gemm_ln(h, self.mlp[2].weight, self.norm2.weight, self.norm2.bias, self.norm2.eps, residual=x)


    def cross_pair(self, both, n: int, frames: int):
        """Inference path of one 'cross' layer over both streams held in one buffer: both[:n] = image tokens,
        both[n:] = point tokens.  Reference order (model/transformer/transformer.py:98-100): the image stream attends to
        the old point stream, then the point stream attends to the UPDATED image stream.  Everything that depends only
        on the layer input -- the query projection (+ its sequence-axis normalisation) and the x half of mlp[0] -- is one
        call over both streams; the rest runs stream after stream, writing into the halves of the output buffer."""
        C = both.shape[1]
        scale = 1.0 / self.dim ** 0.5
        tc = _attn_tc(n)
        q = ops.colnorm_rows(ops.gemm(both, self.q_proj.weight), 2 * frames)
        w1 = self.mlp[0].weight
        h = ops.gemm(both, w1[:, :C])
        out = torch.empty_like(both)
        for lo, src in ((0, both[n:]), (n, out[:n])):
            k = ops.gemm(src, self.k_proj.weight)
            if tc:
                msg = ops.attention_vt(q[lo:lo + n], k, ops.gemm(self.v_proj.weight, src), frames, self.nhead, scale)
            else:
                v = ops.gemm(src, self.v_proj.weight)
                with ops.group("tr_attn"):
                    msg = self.attention(q[lo:lo + n], k, v, frames, self.nhead)
            m = ops.gemm_ln(msg, self.merge.weight, self.norm1.weight, self.norm1.bias, self.norm1.eps)
            hh = ops.gemm(m, w1[:, C:], out=h[lo:lo + n], accumulate=True, act=ops.ACT_RELU)
            ops.gemm_ln(hh, self.mlp[2].weight, self.norm2.weight, self.norm2.bias, self.norm2.eps, residual=both[lo:lo + n],
                        out=out[lo:lo + n])
        return out


class LocalFeatureTransformer(nn.Module):
    def __init__(self, D_MODEL, NHEAD, LAYER_NAMES, ATTENTION):
        super().__init__()
        self.d_model, self.nhead, self.layer_names = D_MODEL, NHEAD, LAYER_NAMES
        layer = LoFTREncoderLayer(D_MODEL, NHEAD, ATTENTION)
        self.layers = nn.ModuleList([copy.deepcopy(layer) for _ in range(len(self.layer_names))])
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward(self, feat0, feat1, frames: int = 1):
        """feat0 [B*L,C] (image tokens), feat1 [B*S,C] (point tokens). 3-D [1,L,C] inputs are accepted too."""
        squeeze = feat0.dim() == 3
        if squeeze:
            assert feat0.shape[0] == 1 and feat1.shape[0] == 1
            feat0, feat1 = feat0[0], feat1[0]
        assert self.d_model == feat0.size(-1), "the feature number of src and transformer must be equal"
        n0 = feat0.shape[0]
        if n0 == feat1.shape[0] and not ad.active(self) and all(nm in ("self", "cross") for nm in self.layer_names):
            # inference with equal token counts: both streams live in one [2n, C] buffer for the whole stack
            both = torch.cat([feat0, feat1], 0)
            for layer, name in zip(self.layers, self.layer_names):
                both = layer(both, None, 2 * frames) if name == "self" else layer.cross_pair(both, n0, frames)
            feat0, feat1 = both[:n0], both[n0:]
            return (feat0.unsqueeze(0), feat1.unsqueeze(0)) if squeeze else (feat0, feat1)
        for layer, name in zip(self.layers, self.layer_names):
            if name == "self" and n0 == feat1.shape[0]:
                # both streams go through the SAME layer weights and a self layer never mixes them, so with equal token
                # counts (1280 super-pixels, 1280 super-points) they are one call over 2*frames segments: every kernel
                # of the layer sees twice the rows (the 80-CTA GEMMs fill the SMs) and the launch count halves
                both = layer(torch.cat([feat0, feat1], 0), None, 2 * frames)
                feat0, feat1 = both[:n0], both[n0:]
            elif name == "self":
                feat0 = layer(feat0, feat0, frames)
                feat1 = layer(feat1, feat1, frames)
            elif name == "cross":
                feat0 = layer(feat0, feat1, frames)
                feat1 = layer(feat1, feat0, frames)  # attends to the already-updated image stream
            else:
                raise KeyError(name)
        if squeeze:
            return feat0.unsqueeze(0), feat1.unsqueeze(0)
        return feat0, feat1
