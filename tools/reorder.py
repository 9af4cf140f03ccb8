"""Spatial re-labelling of a point pyramid (host-side logic on device tensors, plain index algebra).

The reference's dataset hands the model points in random order (`np.random.choice` half-sampling, reference
model/kpconv/preprocess_data.py:58), so the 128 neighbours of consecutive rows are unrelated and every gather kernel
(KPConv aggregate, neighbour max-pool, nearest up-sampling) misses L1 and lives on L2 bandwidth.  Re-labelling every
level in Morton order -- points, features and all three index tables permuted consistently -- leaves every per-point
result unchanged (same rows, other order) and makes neighbouring rows share most of their neighbours.
`morton_permutations` returns, per level, the frame-local permutation `perm` (new row i = old row perm[i]);
`permute_pyramid` applies it; `unpermute_rows` maps per-point results back to the caller's order."""
from typing import Dict, List

import torch

__all__ = ["morton_permutations", "permute_pyramid", "unpermute_rows", "inverse_permutations"]


def _spread10(v: torch.Tensor) -> torch.Tensor:
    v = v & 1023
    v = (v | (v << 16)) & 0x030000FF
    v = (v | (v << 8)) & 0x0300F00F
    v = (v | (v << 4)) & 0x030C30C3
    v = (v | (v << 2)) & 0x09249249
    return v


def morton_permutations(points: List[torch.Tensor], frames: int) -> List[torch.Tensor]:
    """points[l]: [frames*n_l, 3] -> perm[l]: [frames, n_l] int64 (frame-local old row of every new row)."""
    perms = []
    for p in points:
        n = p.shape[0] // frames
        q = p.view(frames, n, 3)
        lo = q.amin(dim=1, keepdim=True)
        ext = (q.amax(dim=1, keepdim=True) - lo).amax(dim=2, keepdim=True).clamp_min(1e-12)
        cell = ((q - lo) * (1023.5 / ext)).clamp_(0, 1023).to(torch.int64)
        code = _spread10(cell[..., 0]) | (_spread10(cell[..., 1]) << 1) | (_spread10(cell[..., 2]) << 2)
        perms.append(torch.argsort(code, dim=1, stable=True))
    return perms


def inverse_permutations(perms: List[torch.Tensor]) -> List[torch.Tensor]:
    """inv[l][f, old] = new, with one extra column n_l -> n_l so that the shadow index maps to itself."""
    out = []
    for p in perms:
        f, n = p.shape
        inv = torch.empty((f, n + 1), dtype=torch.int64, device=p.device)
        inv[:, n] = n
        inv.scatter_(1, p, torch.arange(n, device=p.device).expand(f, n))
        out.append(inv)
    return out


def _rows(t: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    """t: [frames*n, ...] -> rows re-ordered frame by frame."""
    f, n = perm.shape
    g = (perm + torch.arange(f, device=perm.device).view(f, 1) * n).reshape(-1)
    return t.index_select(0, g)


def _table(t: torch.Tensor, perm_q: torch.Tensor, inv_s: torch.Tensor) -> torch.Tensor:
    """index table with rows in a query level and values in a source level: permute the rows, re-label the values."""
    f, nq = perm_q.shape
    rows = _rows(t, perm_q).view(f, nq * t.shape[1])
    return torch.gather(inv_s, 1, rows.clamp_(0, inv_s.shape[1] - 1)).view(f * nq, t.shape[1])


def permute_pyramid(d: Dict, perms: List[torch.Tensor], frames: int) -> Dict:
    inv = inverse_permutations(perms)
    L = len(perms)
    out = dict(d)
    out["points"] = [_rows(d["points"][l], perms[l]).contiguous() for l in range(L)]
    if "feats" in d:
        out["feats"] = _rows(d["feats"], perms[0]).contiguous()
    out["neighbors"] = [_table(d["neighbors"][l], perms[l], inv[l]) for l in range(L)]
    out["subsampling"] = [_table(d["subsampling"][l], perms[l + 1], inv[l]) for l in range(L - 1)]
    out["upsampling"] = [_table(d["upsampling"][l], perms[l], inv[l + 1]) for l in range(L - 1)]
    return out


def unpermute_rows(t: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    """per-point result in the re-labelled order -> the caller's order (row perm[i] of the result = new row i)."""
    f, n = perm.shape
    g = (perm + torch.arange(f, device=perm.device).view(f, 1) * n).reshape(-1)
    out = torch.empty_like(t)
    out.index_copy_(0, g, t)
    return out
