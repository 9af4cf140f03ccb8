"""reference model/kpconv/preprocess_data.py on the B200 kernels: the pyramid + KNN-128 table builder that the
reference's datasets call for every frame right before `CoFiI2P.forward` (data/kitti.py:292, data/nuscenes.py:212).

Same functions, argument meaning and returned dictionary as the reference:

  precompute_point_cloud_stack_mode(points [3,N] ndarray, intensity, normals, lengths, num_stages)   (:36-107)
  precompute_point_cloud_cuda(points, intensity, normals, lengths, num_stages)                        (:145-203)
  knn(nodes [M,3], points [N,3], radius_num) -> [N, radius_num] int64                                 (:131-143)
  square_distance(src [B,N,3], tgt [B,M,3], normalize=False)                                          (:110-129)

What differs: the 13 table searches (open3d.ml KNNSearch on the CPU / a dense N x N distance matrix + topk) are ONE
call of the library's Morton-sorted exact search (csrc/knn.cu), the tables stay on the GPU (set `device='cpu'` to get
host tensors like the reference), rows with equal distances are ordered by index (the reference leaves that order to
nanoflann / torch.topk), and `precompute_pyramid_batch` builds the tables of B frames at once for the batched engine.
The random half-sampling is the reference's own host-side `np.random.choice` (WITH replacement, global numpy RNG), so a
seeded run draws the same pyramid as the reference does.  There is no CPU search path: without the CUDA library these
functions raise."""
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from ... import ops

__all__ = ["precompute_point_cloud_stack_mode", "precompute_point_cloud_cuda", "precompute_pyramid_batch", "knn",
           "square_distance", "half_sample"]

RADIUS_NUM = 128  # reference preprocess_data.py:38,147


def half_sample(points: np.ndarray, num_stages: int, replace: bool = True) -> List[np.ndarray]:
    """[3,N] -> num_stages arrays [N_i,3] fp32, N_i = N >> i (reference :52-68).  replace=True is the reference's
    `np.random.choice(np.arange(n), size=n // 2)`; replace=False is open3d's random_down_sample semantic (:158)."""
    levels = []
    for i in range(num_stages):
        if i > 0:
            n = points.shape[1]
            idx = np.random.choice(np.arange(n), size=n // 2) if replace else np.sort(np.random.permutation(n)[: n // 2])
            points = points[:, idx]
        levels.append(np.ascontiguousarray(points.T, dtype=np.float32))
    return levels


def _device(device) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("preprocess_data: the table builder runs on the GPU (cofii2p_b200 has no CPU path)")
    return torch.device("cuda", torch.cuda.current_device()) if device is None or str(device) == "cpu" else torch.device(device)


def _finish(levels: Sequence[torch.Tensor], tables: Dict, lengths: int, out_device) -> Dict:
    mv = (lambda t: t.to(out_device)) if out_device is not None else (lambda t: t)
    lens, n = [], int(lengths)
    for _ in levels:
        lens.append(n)
        n //= 2
    return {"points": [mv(p) for p in levels], "lengths": lens,
            "neighbors": [mv(t) for t in tables["neighbors"]],
            "subsampling": [mv(t) for t in tables["subsampling"]],
            "upsampling": [mv(t) for t in tables["upsampling"]]}


def precompute_point_cloud_stack_mode(points, intensity, normals, lengths, num_stages, device=None):
    """Reference :36-107.  `points` is the dataset's [3,N] array; `intensity` / `normals` are accepted and unused, as in
    the reference.  Returns {'points','lengths','neighbors','subsampling','upsampling'}; tensors live on the current CUDA
    device unless device='cpu'."""
    dev = _device(device)
    levels = [torch.from_numpy(l).to(dev) for l in half_sample(np.asarray(points), num_stages, replace=True)]
    tables = ops.knn_pyramid(levels, frames=1, k=RADIUS_NUM, mode=ops.KNN_DIRECT)
    return _finish(levels, tables, lengths, "cpu" if str(device) == "cpu" else None)


def precompute_point_cloud_cuda(points, intensity, normals, lengths, num_stages, device=None):
    """Reference :145-203: half-sampling without replacement (open3d random_down_sample; drawn here from numpy's global
    RNG, open3d's private generator cannot be reproduced) and `knn()`'s expanded-form distance."""
    dev = _device(device)
    levels = [torch.from_numpy(l).to(dev) for l in half_sample(np.asarray(points), num_stages, replace=False)]
    tables = ops.knn_pyramid(levels, frames=1, k=RADIUS_NUM, mode=ops.KNN_EXPANDED)
    return _finish(levels, tables, lengths, "cpu" if str(device) == "cpu" else None)


def precompute_pyramid_batch(levels: Sequence[torch.Tensor], frames: int, k: int = RADIUS_NUM, mode: int = ops.KNN_DIRECT,
                             want=("neighbors", "subsampling", "upsampling"), workspace: Optional[torch.Tensor] = None) -> Dict:
    """Tables of B stacked frames (levels[l] = [B*N_l, 3] CUDA tensors, frame-local indices) in two kernel launches:
    the batched entry point next to the reference-shaped single-frame functions."""
    return ops.knn_pyramid(list(levels), frames=frames, k=k, mode=mode, want=want, workspace=workspace)


def square_distance(src, tgt, normalize=False):
    """Reference :110-129 (a dense [B,N,M] matrix; kept for the callers that want the matrix itself)."""
    dist = -2.0 * torch.matmul(src, tgt.permute(0, 2, 1).contiguous())
    if normalize:
        dist = dist + 2
    else:
        dist = dist + torch.sum(src ** 2, dim=-1).unsqueeze(-1)
        dist = dist + torch.sum(tgt ** 2, dim=-1).unsqueeze(-2)
    return torch.clamp(dist, min=1e-12, max=None)


def knn(nodes, points, radius_num):
    """Reference :131-143: for every row of `points`, the `radius_num` nearest rows of `nodes` (expanded-form
    distance), without the [N,M] matrix."""
    dev = _device(nodes.device if torch.is_tensor(nodes) and nodes.is_cuda else None)
    out = ops.knn_table(nodes.to(dev, torch.float32), points.to(dev, torch.float32), frames=1, k=int(radius_num),
                        mode=ops.KNN_EXPANDED)
    return out if nodes.is_cuda else out.to(nodes.device)
