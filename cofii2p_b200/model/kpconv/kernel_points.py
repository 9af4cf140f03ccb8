"""Kernel-point dispositions for rigid KPConv without open3d/matplotlib.

The reference optimises 15 points (one fixed at the centre) by repulsion inside the unit sphere, caches the
result as a PLY next to its sources and then applies, per KPConv instance, a random rotation about z, N(0,0.01)
noise and the radius scale (reference model/kpconv/kernel_points.py:389-455).  Trained checkpoints carry the
resulting `kernel_points` buffers in their state_dict, so this generator only matters for from-scratch
construction.  Here the disposition is computed in memory (deterministic, no file I/O): a Fibonacci sphere
relaxed by a few hundred Coulomb-repulsion steps with a centre attraction, rescaled so that the mean norm of
the non-centre points equals 0.66 (reference :381-383, `ratio`).
"""
import math

import numpy as np

_CACHE = {}

# The relaxed 15-point disposition (centre + 14), frozen to 6 decimals so that every machine builds
# bit-identical kernels (the optimiser below reproduces it; other sizes are optimised on the fly).
_K15 = [[0.0, 0.0, 0.0], [0.145336, -0.39351, 0.510844], [-0.20667, 0.064327, 0.624567],
        [0.433223, 0.224999, 0.435122], [-0.464781, -0.413053, 0.224277], [-0.152734, 0.561331, 0.313869],
        [0.551277, -0.354343, 0.086422], [-0.636872, 0.138225, 0.110553], [0.253207, 0.600498, -0.110553],
        [-0.02718, -0.654772, -0.086422], [-0.371362, 0.447784, -0.313869], [0.605212, 0.142661, -0.224277],
        [-0.433223, -0.224999, -0.435122], [0.066238, 0.206065, -0.624567], [0.238329, -0.345213, -0.510844]]


def _disposition(num_kpoints: int, dimension: int = 3, ratio: float = 0.66) -> np.ndarray:
    key = (num_kpoints, dimension)
    if key in _CACHE:
        return _CACHE[key].copy()
    assert dimension == 3, "only 3-D kernels are used on the hot path"
    if num_kpoints == 15:
        return np.array(_K15, dtype=np.float64)
    n = num_kpoints - 1
    i = np.arange(n) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    theta = math.pi * (1 + 5 ** 0.5) * i
    pts = np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)], 1)
    pts = np.concatenate([np.zeros((1, 3)), pts], 0)
    for _ in range(400):  # repulsion between all points + attraction to the centre, centre fixed
        d = pts[:, None, :] - pts[None, :, :]
        r2 = (d ** 2).sum(-1) + 1e-9
        np.fill_diagonal(r2, np.inf)
        f = (d / (r2 ** 1.5)[..., None]).sum(1) - 2.0 * pts
        f[0] = 0.0
        pts = pts + 0.01 * f / max(1e-9, np.abs(f).max())
    pts *= ratio / np.mean(np.linalg.norm(pts[1:], axis=1))
    _CACHE[key] = pts
    return pts.copy()


def load_kernels(radius, num_kpoints, dimension=3, fixed="center", lloyd=False):
    """Same signature and post-processing as the reference's `load_kernels` (np.random-driven z rotation and
    noise), disposition computed in memory."""
    kp = _disposition(num_kpoints, dimension)
    theta = np.random.rand() * 2 * np.pi
    c, s = np.cos(theta), np.sin(theta)
    R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float32)
    kp = kp + np.random.normal(scale=0.01, size=kp.shape)
    kp = radius * kp
    kp = np.matmul(kp, R)
    return kp.astype(np.float32)
