"""Synthetic KITTI-shaped frames for the CoFiI2P hot path (host logic).

A frame is exactly what the reference's dataset hands to `CoFiI2P.forward`
(reference `data/kitti.py:374-393`, `train.py:192-226`):

  pc_data_dict = {'points': 5 x [N_i,3] f32, 'neighbors': 5 x [N_i,128] i64,
                  'subsampling': 4 x [N_{i+1},128] i64, 'upsampling': 4 x [N_i,128] i64,
                  'feats': [N_0,4] f32 (intensity + unit normal), 'lengths': 5 ints}
  img [1,3,160,512] f32 in [0,1]; fine_center_kpt_coors [2,64] i32 (x,y at 1/2 resolution);
  fine_pc_inline_index [64] i64 (rows of the level-1 cloud)

Reproducibility across machines matters (goldens are produced in the build container and checked on
the GPU box), so the geometry lives on an *integer* 0.1 m voxel lattice and the KNN-128 tables are built
from exact integer squared distances with ties broken by index: every backend (CPU BLAS, CUDA) yields the
same tables bit for bit.  The reference builds its tables with open3d KNNSearch on the float cloud
(reference `model/kpconv/preprocess_data.py:75-99`): same semantics (ascending distance, self first).
The pyramid is successive half-sampling *without* replacement (the reference samples with replacement,
`preprocess_data.py:58`, which creates duplicate points whose mutual order is arbitrary).
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional

import numpy as np
import torch

KNN = 128
LEVELS = 5
IMG_H, IMG_W = 160, 512
NUM_KPT = 64
VOXEL = 0.1


# --------------------------------------------------------------------------------------- geometry
def _scene_lattice(rng: np.random.Generator, want: int) -> np.ndarray:
    """Integer voxel coordinates (int32 [n,3], unique) of a street-like scene in the camera frame
    (x right, y down, z forward): ground plane at y~1.65 m, vertical walls, a few boxes."""
    chunks = []
    # ground: density falls off with range like a spinning LiDAR
    n_g = int(want * 3.0)
    r = rng.uniform(2.5, 70.0, n_g)
    th = rng.uniform(0.0, 2 * math.pi, n_g)
    g = np.stack([r * np.cos(th), np.full(n_g, 1.65) + rng.normal(0, 0.02, n_g), r * np.sin(th)], 1)
    chunks.append(g)
    # walls
    for _ in range(int(rng.integers(6, 11))):
        d = rng.uniform(6.0, 45.0)
        a = rng.uniform(0, 2 * math.pi)
        c = np.array([d * math.cos(a), 0.0, d * math.sin(a)])
        t = a + math.pi / 2 + rng.uniform(-0.6, 0.6)
        u = np.array([math.cos(t), 0.0, math.sin(t)])
        w, h = rng.uniform(6.0, 30.0), rng.uniform(2.0, 6.0)
        n_w = int(want * 0.25)
        s = rng.uniform(-w / 2, w / 2, n_w)
        y = rng.uniform(1.65 - h, 1.65, n_w)
        p = c[None, :] + s[:, None] * u[None, :]
        p[:, 1] = y
        chunks.append(p)
    # boxes (cars)
    for _ in range(int(rng.integers(4, 9))):
        c = np.array([rng.uniform(-25, 25), 0.0, rng.uniform(-25, 35)])
        size = np.array([rng.uniform(1.5, 2.0), rng.uniform(1.3, 1.8), rng.uniform(3.5, 5.0)])
        n_b = int(want * 0.04)
        q = rng.uniform(-0.5, 0.5, (n_b, 3))
        face = rng.integers(0, 3, n_b)
        sign = rng.integers(0, 2, n_b) * 2 - 1
        q[np.arange(n_b), face] = 0.5 * sign
        p = c[None, :] + q * size[None, :]
        p[:, 1] = 1.65 - size[1] / 2 + q[:, 1] * size[1]
        chunks.append(p)
    pts = np.concatenate(chunks, 0)
    ijk = np.round(pts / VOXEL).astype(np.int32)
    ijk = np.unique(ijk, axis=0)
    rng.shuffle(ijk, axis=0)
    return ijk


def _knn_table(src: torch.Tensor, qry: torch.Tensor, k: int, chunk: int = 2048) -> torch.Tensor:
    """Exact KNN on integer lattices. src [Ns,3] int, qry [Nq,3] int -> [Nq,k] int64, ascending
    (distance, index). If Ns < k the tail is the shadow index Ns (reference semantic: index == N means
    'no neighbour', `model/kpconv/kpconv.py:91,105`)."""
    ns = src.shape[0]
    sf = src.to(torch.float32)
    qf = qry.to(torch.float32)
    s2 = (sf * sf).sum(1)
    out = torch.empty((qry.shape[0], k), dtype=torch.int64, device=src.device)
    idx = torch.arange(ns, device=src.device, dtype=torch.int64)
    kk = min(k, ns)
    for a in range(0, qry.shape[0], chunk):
        q = qf[a:a + chunk]
        # all quantities are integers < 2^24: exact in fp32 whatever the summation order
        d2 = (q * q).sum(1, keepdim=True) + s2[None, :] - 2.0 * (q @ sf.t())
        key = d2.to(torch.int64) * ns + idx[None, :]
        top = torch.topk(key, kk, dim=1, largest=False, sorted=True).values
        out[a:a + chunk, :kk] = top % ns
    if kk < k:
        out[:, kk:] = ns
    return out


def _rot_y(yaw: float) -> np.ndarray:
    c, s = math.cos(yaw), math.sin(yaw)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]], dtype=np.float64)


def make_frame(seed: int, num_pc: int = 20480, levels: int = LEVELS, knn: int = KNN,
               device: str = "cpu", num_kpt: int = NUM_KPT, cache_dir: Optional[str] = None) -> Dict:
    """Build one synthetic frame (all tensors on CPU). `device` only selects where the KNN tables are
    computed (results are identical). With `cache_dir`, the integer tables are cached on disk."""
    rng = np.random.default_rng(1000 + seed)
    ijk = _scene_lattice(rng, num_pc)
    while ijk.shape[0] < num_pc:  # pathological tiny scenes: regenerate denser
        ijk = np.unique(np.concatenate([ijk, _scene_lattice(rng, num_pc * 2)], 0), axis=0)
        rng.shuffle(ijk, axis=0)
    ijk = ijk[:num_pc]

    # pyramid: level i+1 = first half of a permutation of level i (sampling without replacement)
    lat: List[np.ndarray] = [ijk]
    for _ in range(1, levels):
        perm = rng.permutation(lat[-1].shape[0])[: lat[-1].shape[0] // 2]
        lat.append(lat[-1][perm])

    # random pose of the reference's augmentation (reference data/options.py:33-38): yaw about y, tx, tz
    yaw = rng.uniform(-2 * math.pi, 2 * math.pi)
    tx, tz = rng.uniform(-10, 10), rng.uniform(-10, 10)
    R = _rot_y(yaw)

    def to_cloud(l: np.ndarray) -> torch.Tensor:
        p = l.astype(np.float64) * VOXEL
        x = R[0, 0] * p[:, 0] + R[0, 1] * p[:, 1] + R[0, 2] * p[:, 2] + tx
        y = R[1, 0] * p[:, 0] + R[1, 1] * p[:, 1] + R[1, 2] * p[:, 2]
        z = R[2, 0] * p[:, 0] + R[2, 1] * p[:, 1] + R[2, 2] * p[:, 2] + tz
        return torch.from_numpy(np.stack([x, y, z], 1).astype(np.float32))

    points = [to_cloud(l) for l in lat]

    cache = None
    if cache_dir is not None:
        os.makedirs(cache_dir, exist_ok=True)
        cache = os.path.join(cache_dir, f"knn_s{seed}_n{num_pc}_l{levels}_k{knn}.npz")
    if cache is not None and os.path.isfile(cache):
        z = np.load(cache)
        neighbors = [torch.from_numpy(z[f"n{i}"].astype(np.int64)) for i in range(levels)]
        subsampling = [torch.from_numpy(z[f"s{i}"].astype(np.int64)) for i in range(levels - 1)]
        upsampling = [torch.from_numpy(z[f"u{i}"].astype(np.int64)) for i in range(levels - 1)]
    else:
        tl = [torch.from_numpy(l).to(device) for l in lat]
        neighbors = [_knn_table(tl[i], tl[i], knn).cpu() for i in range(levels)]
        subsampling = [_knn_table(tl[i], tl[i + 1], knn).cpu() for i in range(levels - 1)]
        upsampling = [_knn_table(tl[i + 1], tl[i], knn).cpu() for i in range(levels - 1)]
        if cache is not None:
            arrs = {}
            for i in range(levels):
                arrs[f"n{i}"] = neighbors[i].numpy().astype(np.int32)
            for i in range(levels - 1):
                arrs[f"s{i}"] = subsampling[i].numpy().astype(np.int32)
                arrs[f"u{i}"] = upsampling[i].numpy().astype(np.int32)
            np.savez(cache + ".tmp.npz", **arrs)
            os.replace(cache + ".tmp.npz", cache)

    # features: intensity U(0,1) + random unit normal (reference data/kitti.py:293)
    inten = rng.uniform(0, 1, (num_pc, 1))
    nrm = rng.normal(size=(num_pc, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    feats = torch.from_numpy(np.concatenate([inten, nrm], 1).astype(np.float32))

    g = torch.Generator().manual_seed(1000 + seed)
    img = torch.rand((1, 3, IMG_H, IMG_W), generator=g, dtype=torch.float32)

    # key points: coarsest-level points whose camera projection lands inside the 1/2-resolution map
    # (mirrors reference data/kitti.py:334-371); scene lattice coordinates ARE the camera frame.
    fx = fy = 180.0
    cx, cy = IMG_W / 4.0, IMG_H / 4.0
    cam = lat[-1].astype(np.float64) * VOXEL
    z = cam[:, 2]
    u = fx * cam[:, 0] / np.maximum(z, 1e-6) + cx
    v = fy * cam[:, 1] / np.maximum(z, 1e-6) + cy
    ok = (z > 1.0) & (u >= 3) & (u < IMG_W / 2 - 3) & (v >= 3) & (v < IMG_H / 2 - 3)
    cand = np.nonzero(ok)[0]
    if cand.shape[0] == 0:
        cand = np.arange(lat[-1].shape[0])
        u = np.full_like(u, IMG_W / 4.0)
        v = np.full_like(v, IMG_H / 4.0)
    sel = cand[rng.permutation(cand.shape[0])[:num_kpt]]
    if sel.shape[0] < num_kpt:
        sel = np.concatenate([sel, rng.choice(cand, num_kpt - sel.shape[0])])
    kpt_xy = np.stack([np.floor(u[sel]), np.floor(v[sel])], 0).astype(np.int32)  # [2,n] (x,y)
    # nearest level-1 point of each key point (reference data/kitti.py:373 uses point2node): level-4 is a
    # subset of level-1 by construction, so the nearest node is the point itself -> exact lattice match
    l1 = {tuple(p): i for i, p in enumerate(lat[1].tolist())}
    inline = np.array([l1[tuple(p)] for p in lat[-1][sel].tolist()], dtype=np.int64)

    P = np.eye(4)
    P[:3, :3] = R
    P[:3, 3] = [tx, 0.0, tz]
    # training supervision of the reference's loader (reference data/kitti.py:334-371, consumed at train.py:204-246):
    # in-frustum super points (= the key points), as many out-of-frustum ones, their 1/8-resolution pixel index,
    # and the camera model at 1/8 resolution with the cloud->camera pose.  Drawn after every other random number.
    out_cand = np.nonzero(~ok)[0]
    if out_cand.shape[0] == 0:
        out_cand = np.arange(lat[-1].shape[0])
    outline = rng.choice(out_cand, num_kpt, replace=out_cand.shape[0] < num_kpt)
    coarse_pix = (np.floor(v[sel] / 4.0) * (IMG_W // 8) + np.floor(u[sel] / 4.0)).astype(np.int64)
    P_cam = np.eye(4)
    P_cam[:3, :3] = R.T
    P_cam[:3, 3] = -R.T @ np.array([tx, 0.0, tz])
    return {
        "pc_kpt_idx": torch.from_numpy(sel.astype(np.int64)),
        "pc_outline_idx": torch.from_numpy(outline.astype(np.int64)),
        "coarse_img_kpt_idx": torch.from_numpy(coarse_pix),
        "K_4": torch.tensor([[fx / 4, 0, cx / 4], [0, fy / 4, cy / 4], [0, 0, 1]], dtype=torch.float32),
        "P": torch.from_numpy(P_cam.astype(np.float32)),
        "pc_data_dict": {
            "points": points,
            "neighbors": neighbors,
            "subsampling": subsampling,
            "upsampling": upsampling,
            "feats": feats,
            "lengths": [int(p.shape[0]) for p in points],
        },
        "img": img,
        "fine_center_kpt_coors": torch.from_numpy(kpt_xy),
        "fine_xy": torch.from_numpy(np.stack([u[sel], v[sel]], 0).astype(np.float32)),
        "fine_pc_inline_index": torch.from_numpy(inline),
        "P_cloud_from_cam": torch.from_numpy(P.astype(np.float32)),
        "K_half": torch.tensor([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=torch.float32),
        "seed": seed,
    }


def frame_to(frame: Dict, device) -> Dict:
    """Move every tensor of a frame to `device` (what reference train.py:192-217 does by hand)."""
    def mv(x):
        if torch.is_tensor(x):
            return x.to(device)
        if isinstance(x, list):
            return [mv(y) for y in x]
        if isinstance(x, dict):
            return {k: mv(v) for k, v in x.items()}
        return x
    return mv(frame)


def stack_frames(frames: List[Dict]) -> Dict:
    """Stack B frames of equal size along the row axis (index tables stay frame-local): the batched input of
    `CoFiI2P.forward_batch` / the inference engine."""
    B = len(frames)
    d0 = frames[0]["pc_data_dict"]
    L = len(d0["points"])
    pc = {
        "points": [torch.cat([f["pc_data_dict"]["points"][i] for f in frames], 0) for i in range(L)],
        "neighbors": [torch.cat([f["pc_data_dict"]["neighbors"][i] for f in frames], 0) for i in range(L)],
        "subsampling": [torch.cat([f["pc_data_dict"]["subsampling"][i] for f in frames], 0) for i in range(L - 1)],
        "upsampling": [torch.cat([f["pc_data_dict"]["upsampling"][i] for f in frames], 0) for i in range(L - 1)],
        "feats": torch.cat([f["pc_data_dict"]["feats"] for f in frames], 0),
        "lengths": d0["lengths"],
    }
    return {
        "frames": B,
        "pc_data_dict": pc,
        "img": torch.cat([f["img"] for f in frames], 0),
        "fine_center_kpt_coors": [f["fine_center_kpt_coors"] for f in frames],
        "fine_pc_inline_index": [f["fine_pc_inline_index"] for f in frames],
        **{k: [f[k] for f in frames] for k in ("fine_xy", "pc_kpt_idx", "pc_outline_idx", "coarse_img_kpt_idx", "K_4", "P")
           if k in frames[0]},
    }
