"""The step right after the hot path: from the model's test-mode outputs to a camera pose,
as reference evaluation/eval_all.py:99-117 does it.

  correspondences  -- :99-105  16-way cosine arg-max of every selected point feature against its 4x4 pixel patch
                      (ops.fine_match, exact) and the pixel coordinate assembly, including the reference's
                      `x += idx // 4, y += idx % 4` convention
  solve_pose       -- :107-113 OpenCV solvePnPRansac (10000 iterations) + Rodrigues: the reference's own call, kept for parity
  solve_pose_gpu / solve_pose_batch -- the same step with the 10000-hypothesis loop on the device (csrc/pnp.cu: batched
                      P3P-RANSAC, one thread per hypothesis, all frames in one launch) and OpenCV's final refinement
                      (solvePnP ITERATIVE on the inliers, what solvePnPRansac ends with) on the host
  pose_error       -- :16-22   get_P_diff: translation norm (RTE, metres) and summed |euler xzy| (RRE, degrees)
"""
from typing import Dict, Tuple

import numpy as np
import torch

from .model.network import fine_match

__all__ = ["correspondences", "solve_pose", "solve_pose_gpu", "solve_pose_batch", "pose_error", "register"]


def correspondences(outputs) -> Tuple[np.ndarray, np.ndarray, torch.Tensor]:
    """outputs: the 8-tuple of CoFiI2P.forward(mode='test').  Returns (imagePoints [n,2], objectPoints [n,3], idx [n])."""
    patch, fine_pc, fine_center_xy, coarse_pc_points = outputs[4], outputs[5], outputs[6], outputs[7]
    idx, fine_xy = fine_match(patch, fine_pc, fine_center_xy)
    return fine_xy.t().contiguous().cpu().numpy(), coarse_pc_points.cpu().numpy(), idx


def solve_pose(K: np.ndarray, image_points: np.ndarray, object_points: np.ndarray, iterations: int = 10000, seed: int = 0):
    """cv2.solvePnPRansac exactly as eval_all.py:107 calls it (the RNG is re-seeded so that a call is reproducible).
    Returns (success, T [4,4], inliers)."""
    import cv2
    cv2.setRNGSeed(seed)
    ok, rvec, t, inliers = cv2.solvePnPRansac(cameraMatrix=np.asarray(K, dtype=np.float64),
                                              imagePoints=np.ascontiguousarray(image_points, dtype=np.float32),
                                              objectPoints=np.ascontiguousarray(object_points, dtype=np.float32),
                                              iterationsCount=iterations, distCoeffs=None)
    T = np.eye(4)
    if ok:
        R, _ = cv2.Rodrigues(rvec)
        T[0:3, 0:3] = R
        T[0:3, 3:] = t
    return bool(ok), T, inliers


def _refine(K, image_points, object_points, inliers, R0, t0):
    """cv2.solvePnP(ITERATIVE) on the inlier set, as cv::solvePnPRansac finishes (modules/calib3d/src/solvepnp.cpp)."""
    import cv2
    T = np.eye(4)
    T[0:3, 0:3], T[0:3, 3] = R0, t0
    if int(inliers.sum()) < 4:
        return False, T
    ok, rvec, t = cv2.solvePnP(np.ascontiguousarray(object_points[inliers], dtype=np.float32),
                               np.ascontiguousarray(image_points[inliers], dtype=np.float32),
                               np.asarray(K, dtype=np.float64), None, flags=cv2.SOLVEPNP_ITERATIVE)
    if ok:
        R, _ = cv2.Rodrigues(rvec)
        T[0:3, 0:3], T[0:3, 3:] = R, t
    return bool(ok), T


def solve_pose_batch(K: np.ndarray, image_points: torch.Tensor, object_points: torch.Tensor, count: torch.Tensor = None,
                     iterations: int = 10000, threshold: float = 8.0, seed: int = 0):
    """Pose of B frames from padded device correspondences (image_points [B,n,2], object_points [B,n,3], count [B] or the
    engine's [B,2] match counts): one batched RANSAC launch, then the host refinement per frame.
    Returns a list of (success, T [4,4], inlier indices)."""
    from . import ops
    B = image_points.shape[0]
    K = np.asarray(K, dtype=np.float64)
    cam = torch.tensor([[K[0, 0], K[1, 1], K[0, 2], K[1, 2]]] * B, dtype=torch.float32, device=image_points.device)
    cnt, hyp, pose, inl = ops.pnp_ransac(image_points, object_points, cam, count, iterations, threshold, seed)
    pose, inl = pose.cpu().numpy(), inl.cpu().numpy().astype(bool)
    ip, op = image_points.cpu().numpy(), object_points.cpu().numpy()
    out = []
    for b in range(B):
        ok, T = _refine(K, ip[b], op[b], inl[b], pose[b, :9].reshape(3, 3), pose[b, 9:])
        out.append((ok, T, np.nonzero(inl[b])[0]))
    return out


def solve_pose_gpu(K: np.ndarray, image_points: np.ndarray, object_points: np.ndarray, iterations: int = 10000,
                   threshold: float = 8.0, seed: int = 0):
    """Drop-in for solve_pose with the hypothesis loop on the GPU.  Returns (success, T [4,4], inliers [m,1] like OpenCV)."""
    ip = torch.from_numpy(np.ascontiguousarray(image_points, dtype=np.float32)).cuda()[None]
    op = torch.from_numpy(np.ascontiguousarray(object_points, dtype=np.float32)).cuda()[None]
    ok, T, inl = solve_pose_batch(K, ip, op, None, iterations, threshold, seed)[0]
    return ok, T, inl.reshape(-1, 1).astype(np.int32)


def pose_error(T_pred: np.ndarray, T_gt: np.ndarray) -> Tuple[float, float]:
    """(RTE, RRE) of eval_all.py:16-22."""
    from scipy.spatial.transform import Rotation
    d = np.dot(np.linalg.inv(T_pred), T_gt)
    rte = float(np.linalg.norm(d[0:3, 3]))
    rre = float(np.sum(np.abs(Rotation.from_matrix(d[0:3, 0:3]).as_euler("xzy", degrees=True))))
    return rte, rre


def register(model, frame: Dict, K: np.ndarray, T_gt: np.ndarray = None, iterations: int = 10000) -> Dict:
    """One frame end to end: forward(mode='test') -> fine matching -> PnP-RANSAC (-> RTE/RRE when T_gt is given)."""
    args = ("pc_data_dict", "img", "fine_center_kpt_coors", "fine_xy", "fine_pc_inline_index")
    with torch.no_grad():
        out = model(*[frame[k] for k in args], "test")
    img_pts, obj_pts, idx = correspondences(out)
    res = {"image_points": img_pts, "object_points": obj_pts, "fine_index": idx, "success": False, "T": np.eye(4)}
    if img_pts.shape[0] >= 4:
        res["success"], res["T"], res["inliers"] = solve_pose(K, img_pts, obj_pts, iterations)
    if T_gt is not None and res["success"]:
        res["rte"], res["rre"] = pose_error(res["T"], T_gt)
    return res
