"""TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of CoFiI2P's coarse-to-fine correspondence hot path.

Plain PyTorch fp32 on CPU tensors, written functionally over a reference-compatible `state_dict`
(the same 430 keys the reference model owns).  Every function cites the reference file:line it restates
(paths relative to the reference repository WHU-USI3DV/CoFiI2P @ ed90edf).  Nothing here is product
code: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs
may import this module, and only as the checker / the timed CPU baseline.  The product path
(`cofii2p_b200.model`) never falls back to it.

PINNING: the reference ships no tests and no golden vectors, so parity is pinned by
running the *reference itself* in the build container (`oracle/ref_shim.py`) on seeded synthetic frames:
`tests/test_oracle_vs_reference.py` compares this restatement with the reference forward tensor by tensor
(bit-exact on CPU; both dispatch the same ATen ops in the same order) and `oracle/make_golden.py` freezes
reference outputs into `tests/golden/` for the GPU box, where `/root/reference` does not exist.

The op order deliberately follows the reference (including the tensors it materialises, e.g. the
(M,128,15,3) differences of KPConv), because this module is also the CPU baseline that `bench.py` times.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# KPConvFPN layer table: (block name, kind, sigma multiplier, strided)
# reference model/kpconv/kp_backbone.py:11-73 (ctor) with init_radius=0.425, init_sigma=0.2
# (reference model/network.py:24). Strided blocks keep the finer stage's sigma.
INIT_SIGMA = 2 * 0.1
KP_LAYERS: Tuple[Tuple[str, str, float, bool], ...] = (
    ("encoder1_1", "conv", 1, False), ("encoder1_2", "res", 1, False),
    ("encoder2_1", "res", 1, True), ("encoder2_2", "res", 2, False), ("encoder2_3", "res", 2, False),
    ("encoder3_1", "res", 2, True), ("encoder3_2", "res", 4, False), ("encoder3_3", "res", 4, False),
    ("encoder4_1", "res", 4, True), ("encoder4_2", "res", 8, False), ("encoder4_3", "res", 8, False),
    ("encoder5_1", "res", 8, True), ("encoder5_2", "res", 16, False), ("encoder5_3", "res", 16, False),
)
GN_GROUPS = 32
LRELU = 0.1


# ------------------------------------------------------------------------------------------ point stream
def kpconv(s_feats, q_points, s_points, nbr, weights, bias, kernel_points, sigma):
    """Rigid KPConv. reference model/kpconv/kpconv.py:79-122.
    s_feats [N,C], q_points [M,3], s_points [N,3], nbr [M,H] i64 (index N = shadow), weights [K,C,Co]."""
    sp = torch.cat([s_points, torch.zeros_like(s_points[:1]) + 1e6], 0)          # :91 shadow point at 1e6
    nb = sp.index_select(0, nbr.reshape(-1)).view(nbr.shape[0], nbr.shape[1], 3)  # :92
    nb = nb - q_points.unsqueeze(1)                                                # :93
    diff = nb.unsqueeze(2) - kernel_points                                         # :96-97 (M,H,K,3)
    sq = torch.sum(diff ** 2, dim=3)                                               # :98
    w = torch.clamp(1 - torch.sqrt(sq) / sigma, min=0.0)                           # :99 linear influence
    w = torch.transpose(w, 1, 2)                                                   # :100 (M,K,H)
    sf = torch.cat((s_feats, torch.zeros_like(s_feats[:1])), 0)                    # :103 shadow feature 0
    nf = sf.index_select(0, nbr.reshape(-1)).view(nbr.shape[0], nbr.shape[1], -1)  # :104 (M,H,C)
    wf = torch.matmul(w, nf)                                                       # :105 (M,K,C)
    wf = wf.permute(1, 0, 2)                                                       # :108
    ko = torch.matmul(wf, weights)                                                 # :109 (K,M,Co)
    out = torch.sum(ko, dim=0)                                                     # :110
    nsum = torch.sum(nf, dim=-1)                                                   # :113
    nnum = torch.sum(torch.gt(nsum, 0.0), dim=-1)                                  # :114
    nnum = torch.max(nnum, torch.ones_like(nnum))                                  # :115
    out = out / nnum.unsqueeze(1)                                                  # :116
    if bias is not None:
        out = out + bias                                                           # :119-120
    return out


def group_norm_rows(x, weight, bias, groups=GN_GROUPS, eps=1e-5):
    """GroupNorm over all rows of a cloud. reference model/kpconv/modules.py:45-49."""
    y = F.group_norm(x.transpose(0, 1).unsqueeze(0), groups, weight, bias, eps)
    return y.squeeze(0).transpose(0, 1)


def unary(sd: SD, p: str, x, relu=True, norm=True):
    """Linear + GroupNorm (+LeakyReLU 0.1). reference model/kpconv/modules.py:89-94 (UnaryBlock),
    :110-112 (LastUnaryBlock, no norm)."""
    x = F.linear(x, sd[p + ".mlp.weight"], sd[p + ".mlp.bias"])
    if norm:
        x = group_norm_rows(x, sd[p + ".norm.norm.weight"], sd[p + ".norm.norm.bias"])
    if relu:
        x = F.leaky_relu(x, LRELU)
    return x


def maxpool(x, nbr):
    """reference model/kpconv/functional.py:53-66."""
    xs = torch.cat((x, torch.zeros_like(x[:1])), 0)
    return xs.index_select(0, nbr.reshape(-1)).view(nbr.shape[0], nbr.shape[1], -1).max(1)[0]


def nearest_upsample(x, up):
    """Only column 0 of the table is read. reference model/kpconv/functional.py:5-21."""
    xs = torch.cat((x, torch.zeros_like(x[:1])), 0)
    return xs.index_select(0, up[:, 0])


def conv_block(sd: SD, p: str, feats, q, s, nbr, sigma):
    """KPConv + GN + LeakyReLU. reference model/kpconv/modules.py:155-159."""
    x = kpconv(feats, q, s, nbr, sd[p + ".KPConv.weights"], sd.get(p + ".KPConv.bias"),
               sd[p + ".KPConv.kernel_points"], sigma)
    x = group_norm_rows(x, sd[p + ".norm.norm.weight"], sd[p + ".norm.norm.bias"])
    return F.leaky_relu(x, LRELU)


def residual_block(sd: SD, p: str, feats, q, s, nbr, sigma, strided):
    """Bottleneck: unary1 -> KPConv/GN/LReLU -> unary2 (+ shortcut) -> LReLU.
    reference model/kpconv/modules.py:222-240 (forward), :194-218 (which sub-blocks exist)."""
    x = unary(sd, p + ".unary1", feats) if (p + ".unary1.mlp.weight") in sd else feats
    x = kpconv(x, q, s, nbr, sd[p + ".KPConv.weights"], sd.get(p + ".KPConv.bias"),
               sd[p + ".KPConv.kernel_points"], sigma)
    x = group_norm_rows(x, sd[p + ".norm_conv.norm.weight"], sd[p + ".norm_conv.norm.bias"])
    x = F.leaky_relu(x, LRELU)
    x = unary(sd, p + ".unary2", x, relu=False)
    sc = maxpool(feats, nbr) if strided else feats
    if (p + ".unary_shortcut.mlp.weight") in sd:
        sc = unary(sd, p + ".unary_shortcut", sc, relu=False)
    return F.leaky_relu(x + sc, LRELU)


def kpconv_fpn(sd: SD, d: Dict, prefix="pc_encoder", taps: Optional[Dict] = None) -> List[torch.Tensor]:
    """5-stage encoder + 3-stage decoder. reference model/kpconv/kp_backbone.py:79-128.
    Returns [latent_s2 (N1x64), latent_s3 (N2x512), latent_s4 (N3x1024), feats_s5 (N4x2048)]."""
    pts, nbrs, subs, ups = d["points"], d["neighbors"], d["subsampling"], d["upsampling"]
    x = d["feats"]
    stage_out = []
    level = 0
    for name, kind, mult, strided in KP_LAYERS:
        p = f"{prefix}.{name}"
        sigma = INIT_SIGMA * mult
        if strided:
            q, s, nb = pts[level + 1], pts[level], subs[level]
        else:
            q, s, nb = pts[level], pts[level], nbrs[level]
        if kind == "conv":
            x = conv_block(sd, p, x, q, s, nb, sigma)
        else:
            x = residual_block(sd, p, x, q, s, nb, sigma, strided)
        if strided:
            level += 1
        if taps is not None:
            taps[name] = x
        if name.endswith("_3") or name == "encoder1_2":
            stage_out.append(x)  # feats_s1..feats_s5
    f1, f2, f3, f4, f5 = stage_out
    l4 = unary(sd, prefix + ".decoder4", torch.cat([nearest_upsample(f5, ups[3]), f4], 1))   # :111-114
    l3 = unary(sd, prefix + ".decoder3", torch.cat([nearest_upsample(l4, ups[2]), f3], 1))   # :116-119
    l2 = unary(sd, prefix + ".decoder2", torch.cat([nearest_upsample(l3, ups[1]), f2], 1),
               relu=False, norm=False)                                                        # :121-124
    if taps is not None:
        taps.update(decoder4=l4, decoder3=l3, decoder2=l2)
    return [l2, l3, l4, f5]


# ------------------------------------------------------------------------------------------ image stream
def _inorm(x):
    """affine-free InstanceNorm2d (per-instance statistics also in eval). reference model/imagenet.py:123."""
    return F.instance_norm(x, eps=1e-5)


def basic_block(sd: SD, p: str, x, stride):
    """reference model/imagenet.py:57-73."""
    idt = x
    o = F.conv2d(x, sd[p + ".conv1.weight"], None, stride, 1)
    o = F.relu(_inorm(o))
    o = _inorm(F.conv2d(o, sd[p + ".conv2.weight"], None, 1, 1))
    if (p + ".downsample.0.weight") in sd:
        idt = _inorm(F.conv2d(x, sd[p + ".downsample.0.weight"], None, stride, 0))
    return F.relu(o + idt)


def resnet34(sd: SD, img, prefix="img_encoder.backbone", run_dead=False):
    """ResNet-34 with InstanceNorm. reference model/imagenet.py:196-217. Returns (s2, s4, s8) and, with
    `run_dead`, also executes layer3/layer4/avgpool, whose outputs the network never uses
    (reference model/network.py:87-89) -- kept so that the CPU baseline does the reference's work."""
    x = F.relu(_inorm(F.conv2d(img, sd[prefix + ".conv1.weight"], None, 2, 3)))
    s2 = x
    x = F.max_pool2d(x, 3, 2, 1)
    for i in range(3):
        x = basic_block(sd, f"{prefix}.layer1.{i}", x, 1)
    s4 = x
    for i in range(4):
        x = basic_block(sd, f"{prefix}.layer2.{i}", x, 2 if i == 0 else 1)
    s8 = x
    if run_dead:
        y = x
        for i in range(6):
            y = basic_block(sd, f"{prefix}.layer3.{i}", y, 2 if i == 0 else 1)
        for i in range(3):
            y = basic_block(sd, f"{prefix}.layer4.{i}", y, 2 if i == 0 else 1)
        F.adaptive_avg_pool2d(y, (1, 1))
    return s2, s4, s8


def _bn(sd: SD, p: str, x, training=False):
    """BatchNorm2d of the decoder. reference model/imagenet.py:381-394. Eval: running statistics."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training, 0.1, 1e-5)


def residual_conv(sd: SD, p: str, x, training=False):
    """3x3+BN+ReLU, 3x3+BN, 3x3 skip conv+BN, add, ReLU. reference model/imagenet.py:398-411."""
    idt = _bn(sd, p + ".conv_skip.1", F.conv2d(x, sd[p + ".conv_skip.0.weight"], None, 1, 1), training)
    o = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"], None, 1, 1), training))
    o = _bn(sd, p + ".bn2", F.conv2d(o, sd[p + ".conv2.weight"], None, 1, 1), training)
    return F.relu(o + idt)


def image_upsample(sd: SD, p: str, x1, x2, training=False):
    """bilinear x2 (align_corners=False) -> cat -> 2 x ResidualConv. reference model/imagenet.py:440-444."""
    x1 = F.interpolate(x1, scale_factor=2, mode="bilinear", align_corners=False)
    x = torch.cat((x1, x2), 1)
    x = residual_conv(sd, p + ".conv.0", x, training)
    return residual_conv(sd, p + ".conv.1", x, training)


# ------------------------------------------------------------------------------------------ transformer
def posenc_sine(xyz, d_model=128, temperature=10000.0):
    """reference model/transformer/position_encoding.py:29-50 (scale 2*pi, interleaved sin/cos, zero pad)."""
    n_dim = xyz.shape[-1]
    npf = d_model // n_dim // 2 * 2
    pad = d_model - npf * n_dim
    dim_t = torch.arange(npf, dtype=torch.float32, device=xyz.device)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="trunc") / npf)
    xyz = xyz * (2 * math.pi)
    pd = xyz.unsqueeze(-1) / dim_t
    e = torch.stack([pd[..., 0::2].sin(), pd[..., 1::2].cos()], dim=-1).reshape(*xyz.shape[:-1], -1)
    return F.pad(e, (0, pad))


def loftr_layer(sd: SD, p: str, x, source, nhead=4):
    """reference model/transformer/transformer.py:43-64 and linear_attention.py:56-79.
    NB F.normalize(q) with default dim=1 normalises over the SEQUENCE axis (transformer.py:53)."""
    bs, dim = x.size(0), x.size(2) // nhead
    q = F.normalize(F.linear(x, sd[p + ".q_proj.weight"]).view(bs, -1, nhead, dim))
    k = F.linear(source, sd[p + ".k_proj.weight"]).view(bs, -1, nhead, dim)
    v = F.linear(source, sd[p + ".v_proj.weight"]).view(bs, -1, nhead, dim)
    qk = torch.einsum("nlhd,nshd->nlsh", q, k)
    a = torch.softmax((1.0 / dim ** 0.5) * qk, dim=2)
    msg = torch.einsum("nlsh,nshd->nlhd", a, v).contiguous()
    msg = F.linear(msg.view(bs, -1, nhead * dim), sd[p + ".merge.weight"])
    msg = F.layer_norm(msg, (nhead * dim,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-5)
    msg = torch.cat([x, msg], dim=2)
    msg = F.linear(F.relu(F.linear(msg, sd[p + ".mlp.0.weight"])), sd[p + ".mlp.2.weight"])
    msg = F.layer_norm(msg, (nhead * dim,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], 1e-5)
    return x + msg


def local_feature_transformer(sd: SD, f0, f1, prefix="transformer", names=("self", "cross") * 4):
    """reference model/transformer/transformer.py:85-103: the SAME layer weights serve both streams and in
    'cross' the point stream attends to the already-updated image stream."""
    for i, name in enumerate(names):
        p = f"{prefix}.layers.{i}"
        if name == "self":
            f0 = loftr_layer(sd, p, f0, f0)
            f1 = loftr_layer(sd, p, f1, f1)
        else:
            f0 = loftr_layer(sd, p, f0, f1)
            f1 = loftr_layer(sd, p, f1, f0)
    return f0, f1


# ------------------------------------------------------------------------------------------ heads / matching
def pc_feature_layer(sd: SD, x, p="pc_feature_layer"):
    """Linear-LN-ReLU-Linear-LN-ReLU-Linear. reference model/network.py:29."""
    x = F.relu(F.layer_norm(F.linear(x, sd[p + ".0.weight"]), (1024,), sd[p + ".1.weight"], sd[p + ".1.bias"]))
    x = F.relu(F.layer_norm(F.linear(x, sd[p + ".3.weight"]), (512,), sd[p + ".4.weight"], sd[p + ".4.bias"]))
    return F.linear(x, sd[p + ".6.weight"])


def score_head_1d(sd: SD, p: str, x):
    """1x1 conv, affine-free InstanceNorm, ReLU (x2), 1x1 conv, sigmoid. reference model/network.py:42."""
    x = F.relu(F.instance_norm(F.conv1d(x, sd[p + ".0.weight"])))
    x = F.relu(F.instance_norm(F.conv1d(x, sd[p + ".3.weight"])))
    return torch.sigmoid(F.conv1d(x, sd[p + ".6.weight"]))


def score_head_2d(sd: SD, p: str, x):
    """reference model/network.py:43."""
    x = F.relu(F.instance_norm(F.conv2d(x, sd[p + ".0.weight"])))
    x = F.relu(F.instance_norm(F.conv2d(x, sd[p + ".3.weight"])))
    return torch.sigmoid(F.conv2d(x, sd[p + ".6.weight"]))


def fine_process(coarse_pc_score, coarse_pc_feature, coarse_img_feature, thrs=0.9):
    """Score-threshold selection, cosine distance to every super-pixel, argmin, border mask.
    reference model/network.py:167-187. Returns (coarse_xy [2,n] f32 (x=col,y=row), pc_inline_index [n] i64)."""
    score = torch.squeeze(coarse_pc_score)
    sel = torch.where(score >= thrs)[0]
    pcf = coarse_pc_feature[:, sel.long()]                                   # [C, n]
    b, c, h, w = coarse_img_feature.shape
    imf = torch.squeeze(coarse_img_feature.reshape(b, c, h * w))             # [C, HW]
    dist = 1 - torch.sum(imf.unsqueeze(-1) * pcf.unsqueeze(-2), dim=0)       # [HW, n]
    idx = torch.argmin(dist, dim=0)
    xs = torch.linspace(0, w - 1, w).view(1, -1).expand(h, w).unsqueeze(0)
    ys = torch.linspace(0, h - 1, h).view(-1, 1).expand(h, w).unsqueeze(0)
    xy = torch.cat((xs, ys), dim=0).reshape(2, h * w).to(idx.device)
    cxy = xy[:, idx]
    m = (cxy[0] >= 2) & (cxy[0] <= 62) & (cxy[1] <= 18) & (cxy[1] >= 2)      # hard-coded for the 20x64 grid
    return cxy[:, m], sel[m]


def extract_patch(feature_map, center_points, size=4):
    """4x4 window [floor(c-2):floor(c+2)) per centre. reference model/network.py:206-226.
    feature_map [B,C,H,W], center_points [2,n] (x,y) -> [n,B,C,4,4]."""
    lt = torch.floor(center_points - size / 2)
    rb = torch.floor(center_points + size / 2)
    out = []
    for i in range(center_points.shape[1]):
        l, t, r, b = int(lt[0, i]), int(lt[1, i]), int(rb[0, i]), int(rb[1, i])
        patch = feature_map[:, :, t:b, l:r]
        assert patch.shape == (feature_map.size(0), feature_map.size(1), 4, 4)
        out.append(patch)
    return torch.stack(out)


def square_distance(src, tgt):
    """|a|^2+|b|^2-2ab with clamp 1e-12, in this accumulation order. reference model/network.py:228-247."""
    dist = -2.0 * torch.matmul(src, tgt.permute(0, 2, 1).contiguous())
    dist += torch.sum(src ** 2, dim=-1).unsqueeze(-1)
    dist += torch.sum(tgt ** 2, dim=-1).unsqueeze(-2)
    return torch.clamp(dist, min=1e-12, max=None)


def point2node(nodes, points):
    """Nearest node of each point (topk k=1 smallest). reference model/network.py:250-264."""
    dist = square_distance(points.unsqueeze(0), nodes.unsqueeze(0))[0]
    return dist.topk(k=1, dim=-1, largest=False)[1].squeeze(-1)


def fine_match(fine_img_feature_patch, fine_pc_inline_feature, fine_center_xy):
    """Caller-side pixel<->point match: cosine similarity with the 16 patch pixels, argmax, pixel coordinate.
    reference evaluation/eval_all.py:99-105. patch [n,64,16], pc [n,64], centres [2,n] -> (idx [n], xy [2,n])."""
    sim = torch.cosine_similarity(fine_img_feature_patch.unsqueeze(-1),
                                  fine_pc_inline_feature.unsqueeze(-1).unsqueeze(-2))
    idx = torch.argmax(torch.squeeze(sim, -1), dim=1)
    # NB the reference adds idx//4 (the patch ROW offset) to x and idx%4 to y (eval_all.py:104-105); mirrored
    x = fine_center_xy[0] - 2 + torch.div(idx, 4, rounding_mode="floor")
    y = fine_center_xy[1] - 2 + idx % 4
    return idx, torch.stack([x, y], 0)


# ------------------------------------------------------------------------------------------ full forward
def forward(sd: SD, pc_data_dict: Dict, img, fine_center_kpt_coors, fine_xy, fine_pc_inline_index, mode: str,
            run_dead: bool = False, taps: Optional[Dict] = None, img_hw: Tuple[int, int] = (160, 512),
            bn_training: bool = False):
    """CoFiI2P.forward. reference model/network.py:74-164. Returns the reference's 8-tuple."""
    pe_h, pe_w = img_hw[0] // 8, img_hw[1] // 8
    pcs = kpconv_fpn(sd, pc_data_dict, taps=taps)
    s2, s4, s8 = resnet34(sd, img, run_dead=run_dead)
    pc_decode_3 = F.normalize(pcs[0], dim=1, p=2)                                    # :82
    pc_mid = F.normalize(pc_feature_layer(sd, pcs[3]), dim=1, p=2)                   # :84
    img_s8 = F.normalize(s8, dim=1, p=2)                                             # :90
    gy, gx = torch.meshgrid(torch.arange(0, pe_h), torch.arange(0, pe_w), indexing="ij")
    img_xy = torch.stack([gy, gx], -1).reshape(1, pe_h * pe_w, 2).to(img.device)      # :104-105 (row, col)
    img_pos = posenc_sine(img_xy)                                                    # :106
    pc_pos = posenc_sine(pc_data_dict["points"][-1].unsqueeze(0))                    # :107
    f_img = img_s8.flatten(2).transpose(1, 2) + img_pos                              # :113
    f_pc = pc_mid.unsqueeze(0) + pc_pos                                              # :114
    f_img, f_pc = local_feature_transformer(sd, f_img, f_pc)                         # :115
    img_mid = f_img.transpose(1, 2).reshape(1, -1, pe_h, pe_w)                       # :117
    pc_fus = f_pc.transpose(1, 2)                                                    # :118
    pc_score = score_head_1d(sd, "pc_score_layer", pc_fus)                           # :123
    img_score = score_head_2d(sd, "img_score_layer", img_mid)                        # :124
    pc_norm = F.normalize(torch.squeeze(pc_fus), dim=0, p=2)                         # :125
    img_norm = F.normalize(img_mid, dim=1, p=2)                                      # :126
    up4 = image_upsample(sd, "img_upsample_1", img_s8, s4, bn_training)              # :129
    up2 = F.normalize(image_upsample(sd, "img_upsample_2", up4, s2, bn_training), dim=1, p=2)   # :130
    if taps is not None:
        taps.update(img_s2=s2, img_s4=s4, img_s8=s8, pc_mid=pc_mid, tr_img=f_img, tr_pc=f_pc,
                    img_up4=up4, img_up2=up2, pc_decode_3=pc_decode_3)
    if mode in ("train", "val"):
        fine_pc = pc_decode_3[fine_pc_inline_index]                                  # :138
        patch = torch.squeeze(extract_patch(up2, fine_center_kpt_coors))             # :141
        fine_center_xy, coarse_pts = None, None
    elif mode == "test":
        sel, thrs = None, 0.9
        while sel is None or sel.numel() < 4:                                        # :148-151
            coarse_xy, sel = fine_process(pc_score, pc_norm, img_norm, thrs=thrs)
            thrs -= 0.02
        coarse_pts = pc_data_dict["points"][-1][sel]                                 # :152
        cidx = point2node(pc_data_dict["points"][1], coarse_pts)                     # :153
        fine_center_xy = coarse_xy * 4                                               # :156
        patch = torch.squeeze(extract_patch(up2, fine_center_xy))                    # :157
        patch = patch.reshape(patch.shape[0], patch.shape[1], -1)                    # :158
        fine_pc = pc_decode_3[torch.squeeze(torch.Tensor(cidx).long())]              # :161
    else:
        raise ValueError(mode)
    return img_norm, pc_norm, img_score, pc_score, patch, fine_pc, fine_center_xy, coarse_pts
