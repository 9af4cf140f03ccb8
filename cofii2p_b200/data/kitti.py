"""Front-end for the reference's on-disk KITTI layout (reference data/kitti.py:21-393), without open3d.

Directory layout (what the reference's preprocessing scripts write and `kitti_pc_img_dataset` reads, :111-141):

    <root>/calib/<seq>/calib.txt                          P0..P3 (3x4 projection rows), Tr (velodyne -> cam0)
    <root>/sequences/<seq>/img_P2|img_P3/%06d.npy         uint8 [H, W, 3] colour image of camera 2 / 3
    <root>/sequences/<seq>/K_P2|K_P3/%06d.npy             3x3 intrinsics of that image
    <root>/sequences/<seq>/pc_npy_with_normal/%06d.npy    float [7, N]: xyz (velodyne frame), intensity, unit normal

`KittiFrames[i]` returns the dictionary of the reference's `__getitem__` (:259-393) with the same keys, shapes and dtypes --
what `train.py:186-226` / `evaluation/eval_all.py:60-96` move to the GPU and hand to `CoFiI2P.forward`:
voxel-grid (0.1 m) + random down-sampling to `num_pc` points, random rigid transform of the cloud, the 5-level pyramid with
its KNN-128 tables, half-resolution crop of the image with the matching intrinsics, and the training supervision (in-frustum
/ out-of-frustum super-points, their 1/8-resolution pixels, 1/2-resolution patch centres, level-1 node of every key point).

What differs from the reference:
  * the pyramid / table builder is the library's (cofii2p_b200.model.kpconv.preprocess_data -> csrc/knn.cu, on the GPU);
    `table_builder` injects another one (tests run the CPU oracle's there);
  * open3d's voxel_down_sample is restated in numpy: one output point per occupied 0.1 m voxel = the mean of its points,
    intensities and normals (the published behaviour of open3d::geometry::PointCloud::VoxelDownSample; voxel origin
    = min bound - voxel/2);
  * the per-index seeding (:261-264) is kept, so item i is the same on every call."""
import math
import os
import random
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

__all__ = ["KittiCalib", "KittiFrames", "voxel_down_sample", "write_synthetic_sequence"]


class KittiCalib:
    """calib.txt reader (reference :21-66): '<key>: 12 floats'.  P* rows give K and the translation of the camera w.r.t.
    cam0 (tx = (m03 - cx tz) / fx, ...); Tr is the velodyne -> cam0 transform."""

    def __init__(self, root: str):
        self.mat: Dict[int, Dict[str, np.ndarray]] = {}
        base = os.path.join(root, "calib")
        for seq in sorted(os.listdir(base)):
            d = self.mat.setdefault(int(seq), {})
            with open(os.path.join(base, seq, "calib.txt")) as f:
                for line in f:
                    if len(line) < 5:
                        continue
                    key = line[0:2]
                    m = np.array(line[4:].split(), dtype=np.float32).reshape(3, 4)
                    P = np.identity(4, dtype=np.float32)
                    if key == "Tr":
                        P[0:3, :] = m
                    else:
                        K = m[0:3, 0:3]
                        d[key + "_K"] = K
                        tz = m[2, 3]
                        P[0:3, 3] = [(m[0, 3] - K[0, 2] * tz) / K[0, 0], (m[1, 3] - K[1, 2] * tz) / K[1, 1], tz]
                    d[key] = P

    def get_matrix(self, seq: int, key: str) -> np.ndarray:
        return self.mat[seq][key]


def voxel_down_sample(pc: np.ndarray, intensity: np.ndarray, sn: np.ndarray, voxel: float):
    """[3,N], [1,N], [3,N] -> the per-voxel means (reference :144-161 through open3d)."""
    origin = pc.min(axis=1, keepdims=True) - voxel * 0.5
    key = np.floor((pc - origin) / voxel).astype(np.int64)
    _, inv, cnt = np.unique(key.T, axis=0, return_inverse=True, return_counts=True)
    inv = inv.reshape(-1)
    allv = np.concatenate([pc, intensity, sn], 0).astype(np.float64)          # [7, N]
    acc = np.zeros((7, cnt.shape[0]), dtype=np.float64)
    for r in range(7):
        acc[r] = np.bincount(inv, weights=allv[r], minlength=cnt.shape[0])
    acc /= cnt[None, :]
    return acc[0:3].astype(np.float32), acc[3:4].astype(np.float32), acc[4:7].astype(np.float32)


def _rotation(angles) -> np.ndarray:   # reference :205-217, R = Rz Ry Rx
    cx, sx, cy, sy, cz, sz = (math.cos(angles[0]), math.sin(angles[0]), math.cos(angles[1]), math.sin(angles[1]),
                              math.cos(angles[2]), math.sin(angles[2]))
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


class KittiFrames(torch.utils.data.Dataset):
    """opt: the reference's Options_KITTI attributes that matter here -- data_path, num_pc, num_kpt, img_H, img_W and the
    pose-perturbation amplitudes P_t{x,y,z}_amplitude / P_R{x,y,z}_amplitude (reference data/options.py:17-38)."""

    def __init__(self, opt, mode: str, table_builder: Optional[Callable] = None, point2node: Optional[Callable] = None,
                 augment: bool = True, seqs: Optional[List[int]] = None):
        self.opt, self.mode, self.augment = opt, mode, augment
        self.root = opt.data_path
        self.calib = KittiCalib(self.root)
        self.table_builder, self.point2node = table_builder, point2node
        if seqs is None:
            if mode == "train":
                seqs = list(range(9))
            elif mode == "val":
                seqs = [9, 10]
            else:
                raise ValueError("mode must be 'train' or 'val' (reference :119-124)")
        self.items = []
        for seq in seqs:
            sdir = os.path.join(self.root, "sequences", "%02d" % seq)
            if not os.path.isdir(sdir):
                continue
            n = len(os.listdir(os.path.join(sdir, "img_P2")))
            for i in range(n):
                for cam in ("P2", "P3"):
                    self.items.append((os.path.join(sdir, "img_" + cam), os.path.join(sdir, "pc_npy_with_normal"),
                                       os.path.join(sdir, "K_" + cam), seq, i, cam))

    def __len__(self):
        return len(self.items)

    # ---- pieces of the reference's pipeline ----------------------------------------------------------------------
    def _downsample(self, pc, intensity, sn):   # reference :163-175
        n, want = pc.shape[1], self.opt.num_pc
        if n >= want:
            idx = np.random.choice(n, want, replace=False)
        else:
            fix = np.arange(n)
            while n + fix.shape[0] < want:
                fix = np.concatenate((fix, np.arange(n)))
            idx = np.concatenate((fix, np.random.choice(n, want - fix.shape[0], replace=False)))
        return pc[:, idx], intensity[:, idx], sn[:, idx]

    def _random_transform(self) -> np.ndarray:   # reference :219-238
        o = self.opt
        t = [random.uniform(-o.P_tx_amplitude, o.P_tx_amplitude), random.uniform(-o.P_ty_amplitude, o.P_ty_amplitude),
             random.uniform(-o.P_tz_amplitude, o.P_tz_amplitude)]
        a = [random.uniform(-o.P_Rx_amplitude, o.P_Rx_amplitude), random.uniform(-o.P_Ry_amplitude, o.P_Ry_amplitude),
             random.uniform(-o.P_Rz_amplitude, o.P_Rz_amplitude)]
        P = np.identity(4, dtype=np.float32)
        P[0:3, 0:3] = _rotation(a)
        P[0:3, 3] = t
        return P

    def _pyramid(self, pc, intensity, sn) -> Dict:
        if self.table_builder is not None:
            return self.table_builder(pc, intensity, sn, self.opt.num_pc, 5)
        from ..model.kpconv.preprocess_data import precompute_point_cloud_stack_mode
        return precompute_point_cloud_stack_mode(pc, intensity, sn, lengths=self.opt.num_pc, num_stages=5)

    def __getitem__(self, index: int) -> Dict:
        import cv2
        o = self.opt
        (seed,) = np.random.SeedSequence([index]).generate_state(1)     # reference :261-264
        np.random.seed(int(seed))
        random.seed(int(seed))
        img_dir, pc_dir, k_dir, seq, i, cam = self.items[index]
        img = np.load(os.path.join(img_dir, "%06d.npy" % i))
        data = np.load(os.path.join(pc_dir, "%06d.npy" % i))
        K = np.load(os.path.join(k_dir, "%06d.npy" % i))
        pc, intensity, sn = data[0:3], data[3:4], data[4:]
        P_Tr = self.calib.get_matrix(seq, cam) @ self.calib.get_matrix(seq, "Tr")
        pc = P_Tr[0:3, 0:3] @ pc + P_Tr[0:3, 3:]                         # into the camera frame (:273-275)
        sn = P_Tr[0:3, 0:3] @ sn
        pc, intensity, sn = voxel_down_sample(pc, intensity, sn, 0.1)   # :280
        pc, intensity, sn = self._downsample(pc, intensity, sn)         # :281
        P = self._random_transform()                                    # :283-285
        pc = (P[0:3, 0:3] @ pc + P[0:3, 3:]).astype(np.float32)
        sn = (P[0:3, 0:3] @ sn).astype(np.float32)
        d = self._pyramid(pc, intensity, sn)                            # :289
        d["feats"] = torch.from_numpy(np.concatenate([intensity, sn], 0).T.astype(np.float32))
        for l in range(5):
            d["neighbors"][l] = d["neighbors"][l].long()
            if l < 4:
                d["subsampling"][l], d["upsampling"][l] = d["subsampling"][l].long(), d["upsampling"][l].long()
        coarse = d["points"][-1].detach().cpu().numpy().astype(np.float32).T           # [3, n4]
        # image: half resolution, crop to img_H x img_W, intrinsics follow (:304-319)
        img = cv2.resize(img, (int(round(img.shape[1] * 0.5)), int(round(img.shape[0] * 0.5))), interpolation=cv2.INTER_LINEAR)
        K = 0.5 * K
        K[2, 2] = 1
        if self.mode == "train":
            dx, dy = random.randint(0, img.shape[1] - o.img_W), random.randint(0, img.shape[0] - o.img_H)
        else:
            dx, dy = int((img.shape[1] - o.img_W) / 2), int((img.shape[0] - o.img_H) / 2)
        img = img[dy:dy + o.img_H, dx:dx + o.img_W, :]
        K = K.copy()
        K[0, 2] -= dx
        K[1, 2] -= dy
        K_2, K_4 = 0.5 * K, 0.125 * K
        K_2[2, 2] = K_4[2, 2] = 1
        if self.mode == "train" and self.augment:
            from PIL import Image
            from torchvision import transforms
            img = np.array(transforms.ColorJitter((0.8, 1.2), (0.8, 1.2), (0.8, 1.2), (-0.1, 0.1))(Image.fromarray(img)))
        # supervision (:332-371): undo the random transform, project the super-points to 1/8 resolution
        Rinv = np.linalg.inv(P[0:3, 0:3])
        cam_pts = Rinv @ coarse - Rinv @ P[0:3, 3:]
        proj = K_4 @ cam_pts
        proj[0:2] = proj[0:2] / proj[2:]
        xy = np.floor(proj[0:2] + 0.5)
        W8, H8 = o.img_W * 0.125, o.img_H * 0.125
        inside = (xy[0] >= 1) & (xy[0] <= W8 - 3) & (xy[1] >= 1) & (xy[1] <= H8 - 3) & (proj[2] > 0)
        in_idx, out_idx = np.where(inside)[0], np.where(~inside)[0]
        pc_kpt_idx = in_idx[np.random.permutation(len(in_idx))[:o.num_kpt]]
        pc_outline_idx = out_idx[np.random.permutation(len(out_idx))[:o.num_kpt]]
        mask8 = np.zeros((int(H8), int(W8)), dtype=np.float32)
        mask8[xy[1, inside].astype(np.int64), xy[0, inside].astype(np.int64)] = 1.0
        coarse_xy = xy[:, pc_kpt_idx]
        img_kpt_idx = xy[1, pc_kpt_idx] * W8 + xy[0, pc_kpt_idx]
        free = np.where(mask8.reshape(-1) == 0)[0]
        img_outline_idx = free[np.random.permutation(len(free))[:o.num_kpt]]
        proj2 = K_2 @ cam_pts[:, pc_kpt_idx]
        proj2[0:2] = proj2[0:2] / proj2[2:]
        fine_xy = np.floor(proj2[0:2])
        ok = (fine_xy[0] >= 0) & (fine_xy[0] <= o.img_W * 0.5 - 1) & (fine_xy[1] >= 0) & (fine_xy[1] <= o.img_H * 0.5 - 1) & (proj2[2] > 0)
        assert np.all(ok), "a key point projects outside the half-resolution image (reference :364)"
        nodes, kpts = d["points"][1], d["points"][-1][torch.from_numpy(pc_kpt_idx).long()]
        if self.point2node is not None:
            inline = self.point2node(nodes, kpts)
        else:
            from ..model.network import point2node
            dev = nodes.device if nodes.is_cuda else torch.device("cuda")
            inline = point2node(nodes.to(dev), kpts.to(dev)).to(nodes.device)
        return {"img": torch.from_numpy(img.astype(np.float32) / 255.0).permute(2, 0, 1).contiguous(),
                "pc_data_dict": d, "fine_pc_inline_index": inline.long(),
                "K": torch.from_numpy(K_2.astype(np.float32)), "K_4": torch.from_numpy(K_4.astype(np.float32)),
                "P": torch.from_numpy(np.linalg.inv(P).astype(np.float32)), "index": index,
                "coarse_img_mask": torch.from_numpy(mask8).float(),
                "pc_kpt_idx": torch.from_numpy(pc_kpt_idx), "pc_outline_idx": torch.from_numpy(pc_outline_idx),
                "fine_xy_coors": torch.from_numpy(fine_xy.astype(np.int32)),
                "coarse_img_kpt_idx": torch.from_numpy(img_kpt_idx).long(),
                "fine_img_kpt_index": torch.from_numpy(fine_xy[1] * o.img_W * 0.5 + fine_xy[0]).long(),
                "fine_center_kpt_coors": torch.from_numpy((coarse_xy * 4).astype(np.int32)),
                "coarse_img_outline_index": torch.from_numpy(img_outline_idx).long()}


def write_synthetic_sequence(root: str, seq: int, frames: int, seed: int = 0, n_points: int = 60000,
                             img_hw=(370, 1226)) -> None:
    """Write `frames` synthetic frames in the reference's on-disk layout (tests / dry runs without the KITTI download):
    a street-like velodyne cloud (x forward, y left, z up) with intensities and normals, two colour images, KITTI-like
    calibration."""
    rng = np.random.default_rng(seed)
    sdir = os.path.join(root, "sequences", "%02d" % seq)
    for sub in ("img_P2", "img_P3", "K_P2", "K_P3", "pc_npy_with_normal"):
        os.makedirs(os.path.join(sdir, sub), exist_ok=True)
    os.makedirs(os.path.join(root, "calib", "%02d" % seq), exist_ok=True)
    fx, cx, cy = 718.856, 607.19, 185.22
    K = np.array([[fx, 0, cx], [0, fx, cy], [0, 0, 1]], dtype=np.float32)
    with open(os.path.join(root, "calib", "%02d" % seq, "calib.txt"), "w") as f:
        for key, tx in (("P0", 0.0), ("P1", -386.1448), ("P2", 45.38225), ("P3", -337.2877)):
            f.write("%s: %s\n" % (key, " ".join("%.6e" % v for v in [fx, 0, cx, tx, 0, fx, cy, 0, 0, 0, 1, 0])))
        Tr = [4.276802e-04, -9.999672e-01, -8.084491e-03, -1.198459e-02, -7.210626e-03, 8.081198e-03, -9.999413e-01,
              -5.403984e-02, 9.999738e-01, 4.859485e-04, -7.206933e-03, -2.921968e-01]
        f.write("Tr: %s\n" % " ".join("%.6e" % v for v in Tr))
    for i in range(frames):
        n_g = n_points * 2 // 3
        r, th = rng.uniform(3.0, 60.0, n_g), rng.uniform(-math.pi, math.pi, n_g)
        ground = np.stack([r * np.cos(th), r * np.sin(th), np.full(n_g, -1.7) + rng.normal(0, 0.02, n_g)], 0)
        n_w = n_points - n_g
        side = rng.choice([-1.0, 1.0], n_w)
        wall = np.stack([rng.uniform(2.0, 60.0, n_w), side * rng.uniform(5.0, 9.0, n_w), rng.uniform(-1.7, 2.5, n_w)], 0)
        pc = np.concatenate([ground, wall], 1)
        normals = np.concatenate([np.tile(np.array([[0.0], [0.0], [1.0]]), (1, n_g)),
                                  np.stack([np.zeros(n_w), -side, np.zeros(n_w)], 0)], 1)
        inten = rng.uniform(0.0, 1.0, (1, n_points))
        np.save(os.path.join(sdir, "pc_npy_with_normal", "%06d.npy" % i), np.concatenate([pc, inten, normals], 0).astype(np.float32))
        for cam in ("P2", "P3"):
            np.save(os.path.join(sdir, "img_" + cam, "%06d.npy" % i), rng.integers(0, 256, (img_hw[0], img_hw[1], 3), dtype=np.uint8))
            np.save(os.path.join(sdir, "K_" + cam, "%06d.npy" % i), K)
