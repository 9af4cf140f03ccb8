"""Training losses of the reference's surface (`from model.loss import *`, reference train.py:17;
definitions reference model/loss.py:9-93).

CUDA tensors run the fused forward + analytic-backward kernels of csrc/loss.cu (one launch per loss and direction; the
training step calls them batched over all frames, cofii2p_b200/train.py::fused_training_losses).  The `*_algebra`
functions are the same formulas as plain tensor algebra, written from the reference's definitions and checked bit for bit
against the reference in tests/test_cpu.py: they are what the fused kernels are tested against, and what a CPU tensor gets
(the reference's loss functions are device-agnostic; nothing on the model's forward path has a CPU route)."""
import torch
import torch.nn.functional as F

from .. import autograd as ad
from .. import ops

__all__ = ["desc_loss", "overlap_loss", "fine_circle_loss"]


def desc_loss(device, img_features, pc_features, mask, pos_margin=0.1, neg_margin=1.4, log_scale=10, num_kpt=512):
    """Circle-style log-sum-exp loss over the coarse descriptor distance matrix (reference model/loss.py:69-93).
    img_features [C,N], pc_features [C,M], mask [N,M] (1 = positive pair). Returns (loss, distances)."""
    if img_features.is_cuda and img_features.shape[1] == pc_features.shape[1] and img_features.shape[1] <= 128:
        n = img_features.shape[1]
        img_tok, pc_tok = img_features.t().contiguous(), pc_features.t().contiguous()
        ar = torch.arange(n, device=img_features.device)
        loss = ad.desc_loss_tokens(img_tok, pc_tok, ar, ar, mask.reshape(1, n, n).to(torch.float32), 1, pos_margin, neg_margin,
                                   float(log_scale))
        with torch.no_grad():
            dists = ops.desc_loss_fwd(img_tok, ar, pc_tok, ar, mask.reshape(1, n, n).to(torch.float32), 1, pos_margin,
                                      neg_margin, float(log_scale), want_grad=False, want_dists=True)[1][0]
        return loss[0], dists
    return desc_loss_algebra(device, img_features, pc_features, mask, pos_margin, neg_margin, log_scale, num_kpt)


def desc_loss_algebra(device, img_features, pc_features, mask, pos_margin=0.1, neg_margin=1.4, log_scale=10, num_kpt=512):
    dists = 1 - torch.sum(img_features.unsqueeze(-1) * pc_features.unsqueeze(-2), dim=0)
    neg_mask = 1 - mask
    pos = dists - 1e5 * neg_mask
    pos_w = torch.max(torch.zeros_like(pos), (pos - pos_margin).detach())
    pos_term = log_scale * (pos - pos_margin) * pos_w
    lse_pos_row, lse_pos_col = torch.logsumexp(pos_term, dim=-1), torch.logsumexp(pos_term, dim=-2)
    neg = dists + 1e5 * mask
    neg_w = torch.max(torch.zeros_like(neg), (neg_margin - neg).detach())
    neg_term = log_scale * (neg_margin - neg) * neg_w
    lse_neg_row, lse_neg_col = torch.logsumexp(neg_term, dim=-1), torch.logsumexp(neg_term, dim=-2)
    loss = F.softplus(lse_pos_row + lse_neg_row) / log_scale + F.softplus(lse_pos_col + lse_neg_col) / log_scale
    return torch.mean(loss), dists


def overlap_loss(device, inline_pc_score, outline_pc_score):
    """BCE of the super-point overlap scores: in-frustum -> 1, out-of-frustum -> 0 (reference model/loss.py:53-60)."""
    if inline_pc_score.is_cuda:
        score = torch.cat((inline_pc_score.reshape(-1), outline_pc_score.reshape(-1)), 0)
        idx = torch.arange(score.numel(), device=score.device)
        return ad.overlap_loss_tokens(score, idx, inline_pc_score.numel(), 1)[0]
    return overlap_loss_algebra(device, inline_pc_score, outline_pc_score)


def overlap_loss_algebra(device, inline_pc_score, outline_pc_score):
    score = torch.cat((inline_pc_score, outline_pc_score), 0)
    label = torch.cat((torch.ones(inline_pc_score.shape[0], device=score.device),
                       torch.zeros(outline_pc_score.shape[0], device=score.device)), 0)
    return F.binary_cross_entropy(score, label)


def fine_circle_loss(device, fine_img_feature, fine_pc_feature, relative_index, num_kpt=64):
    """Circle loss (m = 0.2, gamma = 5) between each key point feature and its 4x4 pixel patch
    (reference model/loss.py:9-51). fine_img_feature [n,C,4,4], fine_pc_feature [n,C], relative_index [n] in 0..15."""
    if fine_img_feature.is_cuda and fine_img_feature.shape[1] <= 128 and fine_img_feature.shape[1] % 2 == 0:
        n, C = fine_pc_feature.shape
        return ad.fine_circle_loss_rows(fine_img_feature.reshape(n, C, 16), fine_pc_feature, relative_index.to(torch.int64), 1)[0]
    return fine_circle_loss_algebra(device, fine_img_feature, fine_pc_feature, relative_index, num_kpt)


def fine_circle_loss_algebra(device, fine_img_feature, fine_pc_feature, relative_index, num_kpt=64):
    m, gamma = 0.2, 5
    patch = fine_img_feature.reshape(fine_img_feature.shape[0], fine_img_feature.shape[1], -1)
    sim = torch.cosine_similarity(patch.unsqueeze(-1), fine_pc_feature.unsqueeze(-1).unsqueeze(-2))
    sim = torch.squeeze(sim)
    pos = torch.zeros(num_kpt, 16, device=sim.device)
    pos.scatter_(1, relative_index.view(-1, 1), 1.0)   # pos[arange, relative_index] = 1 without a host-side scalar copy
    neg = 1 - pos
    sp, sn = sim * pos, sim * neg
    ap = torch.relu(-sp.detach() + pos + pos * m)
    an = torch.relu(sn.detach() + neg * m)
    logit_p = -ap * (sp - pos * (1 - m)) * gamma
    logit_n = an * (sn - neg * m) * gamma
    loss_p = torch.sum(torch.exp(logit_p) * pos, dim=-1)
    loss_n = torch.sum(torch.exp(logit_n) * neg, dim=-1)
    return torch.mean(torch.log(1 + loss_n * loss_p))
