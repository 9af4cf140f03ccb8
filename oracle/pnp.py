"""TEST INFRASTRUCTURE ONLY -- numpy (float64) restatement of the device-side P3P-RANSAC (csrc/pnp.cu), the batched
replacement for the hypothesis loop of `cv2.solvePnPRansac(..., iterationsCount=10000)` in reference
evaluation/eval_all.py:107.

OpenCV is a third-party dependency of the reference (unpinned: README.md:35-41 names no version; 4.13 in this image).  Its
RANSAC (modules/calib3d/src/solvepnp.cpp, ptsetreg.cpp) draws minimal sets from cv::RNG, which a device kernel cannot
reproduce, so the algorithm is restated here with a counter-based draw and checked against cv2.solvePnPRansac itself on
synthetic correspondences with a known pose (tests/test_cpu.py::test_pnp_ransac_oracle_vs_opencv): same inlier set, same
pose after the same final refinement (cv2.solvePnP ITERATIVE on the inliers, which is what solvePnPRansac ends with).

  hypothesis h:  4 indices i_k = (Philox4x32-10(h, k, 0, 0; seed)[0] * n) >> 32;  repeated index -> hypothesis skipped
  minimal solve: P3P on the first three (Grunert's distance formulation: s_i^2 + s_j^2 - 2 s_i s_j cos(ij) = d_ij^2 with
                 s2 = u s1, s3 = v s1; eliminating u gives a quartic in v -- coefficients below, derived with sympy --
                 every positive real root gives the three depths, the pose follows from aligning the two triangles)
  disambiguation: the root with the smallest reprojection error of the 4th point
  score:         number of points with squared reprojection error <= threshold^2 (and positive depth)
  winner:        largest score, lowest hypothesis index on ties

Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this module."""
import numpy as np

from .knn import philox4x32_10_word0


def quartic_coeffs(c12, c13, c23, a, b, c):
    """(q0..q4) of the quartic in v, and the pieces of u = -M(v) / L(v); a, b, c = d12^2, d13^2, d23^2."""
    x0 = 2 * c
    x1 = b * x0
    x2 = a * b
    x3 = 2 * x2
    x4 = c23 ** 2
    x5 = 4 * x2
    x6 = x4 * x5
    x7, x8, x9 = a ** 2, b ** 2, c ** 2
    x10 = -a * x0 + x7 + x8 + x9
    x11 = 4 * c13
    x12 = x11 * x2
    x13 = b * c
    x14 = x11 * x13
    x15 = 8 * c13
    x16 = x15 * x2
    x17 = c12 * c23
    x18 = a * c
    x19 = 8 * x18
    x20 = 4 * x13
    x21 = c13 * x19 - x11 * x7 - x11 * x9 + x17 * x20 + x17 * x5 - 4 * x17 * x8
    x22 = c13 ** 2
    x23 = c12 ** 2
    x24 = x20 * x23
    x25 = x13 * x15
    q4 = -x1 + x10 + x3 - x6
    q3 = -x12 + x14 + x16 * x4 + x21
    q2 = (-x16 * x17 - x17 * x25 - 4 * x18 - x19 * x22 + 4 * x22 * x7 + 4 * x22 * x9 + 4 * x23 * x8 - x24 + 4 * x4 * x8 - x6
          + 2 * x7 - 2 * x8 + 2 * x9)
    q1 = x12 - x14 + x21 + x23 * x25
    q0 = x1 + x10 - x24 - x3
    return q0, q1, q2, q3, q4


def u_of_v(v, c12, c13, c23, a, b, c):
    L = 2 * b * c * (c23 * v - c12)
    M = c * (-a * v * v + 2 * a * c13 * v - a - b * v * v + b + c * v * v + c - 2 * c * c13 * v)
    return -M / L if abs(L) > 1e-300 else np.nan


def align(P, X):
    """Rigid (R, t) with R P_i + t = X_i for two congruent triangles P, X [3,3] (rows = points)."""
    def frame(T):
        e1 = T[1] - T[0]
        e1 = e1 / np.linalg.norm(e1)
        e3 = np.cross(e1, T[2] - T[0])
        e3 = e3 / np.linalg.norm(e3)
        return np.stack([e1, np.cross(e3, e1), e3], 1)   # columns
    R = frame(X) @ frame(P).T
    return R, X[0] - R @ P[0]


def p3p(f, P):
    """f [3,3] unit bearing vectors (rows), P [3,3] world points (rows) -> list of (R, t) with X_cam = R X + t."""
    c12, c13, c23 = float(f[0] @ f[1]), float(f[0] @ f[2]), float(f[1] @ f[2])
    a, b, c = (float(np.sum((P[0] - P[1]) ** 2)), float(np.sum((P[0] - P[2]) ** 2)), float(np.sum((P[1] - P[2]) ** 2)))
    if min(a, b, c) < 1e-12:
        return []
    q = quartic_coeffs(c12, c13, c23, a, b, c)
    roots = np.roots(q[::-1]) if abs(q[4]) > 0 else np.roots(q[3::-1])
    out = []
    for r in roots:
        if abs(r.imag) > 1e-7 * max(1.0, abs(r.real)) or r.real <= 0:
            continue
        v = float(r.real)
        u = u_of_v(v, c12, c13, c23, a, b, c)
        if not np.isfinite(u) or u <= 0:
            continue
        den = u * u + v * v - 2 * u * v * c23
        if den <= 0:
            continue
        s1 = np.sqrt(c / den)
        X = np.stack([s1 * f[0], u * s1 * f[1], v * s1 * f[2]], 0)
        # the eliminated system admits spurious roots: keep those that satisfy all three distance equations
        if abs(np.sum((X[0] - X[1]) ** 2) - a) > 1e-6 * a or abs(np.sum((X[0] - X[2]) ** 2) - b) > 1e-6 * b:
            continue
        out.append(align(P, X))
    return out


def reproject(K, R, t, obj):
    Xc = obj @ R.T + t
    z = Xc[:, 2]
    uv = (Xc[:, :2] / z[:, None]) * np.array([K[0, 0], K[1, 1]]) + np.array([K[0, 2], K[1, 2]])
    return uv, z


def sample4(h, n, seed):
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    u = philox4x32_10_word0(np.uint32(h), np.arange(4, dtype=np.uint32), np.uint32(0), np.uint32(0), k0, k1)
    return ((u.astype(np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


def ransac(K, img, obj, iterations=10000, threshold=8.0, seed=0):
    """-> dict(count, hypothesis, R, t, inliers [n] bool) of the winning hypothesis (count 0: none valid)."""
    K, img, obj = np.asarray(K, np.float64), np.asarray(img, np.float64), np.asarray(obj, np.float64)
    n = img.shape[0]
    Kinv = np.linalg.inv(K)
    bear = np.concatenate([img, np.ones((n, 1))], 1) @ Kinv.T
    bear = bear / np.linalg.norm(bear, axis=1, keepdims=True)
    best = dict(count=0, hypothesis=-1, R=np.eye(3), t=np.zeros(3), inliers=np.zeros(n, bool))
    thr2 = threshold * threshold
    for h in range(iterations):
        idx = sample4(h, n, seed)
        if len(set(idx.tolist())) < 4:
            continue
        sols = p3p(bear[idx[:3]], obj[idx[:3]])
        pick, pick_err = None, np.inf
        for R, t in sols:
            uv, z = reproject(K, R, t, obj[idx[3:4]])
            e = float(np.sum((uv[0] - img[idx[3]]) ** 2)) if z[0] > 0 else np.inf
            if e < pick_err:
                pick, pick_err = (R, t), e
        if pick is None:
            continue
        uv, z = reproject(K, pick[0], pick[1], obj)
        inl = (np.sum((uv - img) ** 2, 1) <= thr2) & (z > 0)
        cnt = int(inl.sum())
        if cnt > best["count"]:
            best = dict(count=cnt, hypothesis=h, R=pick[0], t=pick[1], inliers=inl)
    return best
