"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, fp32) of the reference's pyramid/KNN table builder.

Follows reference model/kpconv/preprocess_data.py:
  * half_sample_pyramid  -- :52-68   (`np.random.choice(np.arange(n), size=n // 2)`: WITH replacement, global numpy RNG)
  * knn_table(mode=EXPANDED) -- :110-143 (`square_distance`: -2ab, += |a|^2, += |b|^2, clamp 1e-12; `topk(largest=False)`)
  * knn_table(mode=DIRECT)   -- :75-99   (open3d.ml KNNSearch, an un-vendored dependency absent from this image and
                                          unpinned by the reference -- README.md:35-41 names no version: its published
                                          contract is "k nearest by squared Euclidean distance, ascending", restated as
                                          ((dx*dx + dy*dy) + dz*dz) in fp32)
  * pyramid_tables       -- :75-99 / :172-190 (which cloud queries which: neighbors = level on itself; subsampling =
                                          level i+1 looks up level i; upsampling = level i looks up level i+1)
Ties: torch.topk / nanoflann leave the order of equal distances unspecified; this oracle (and the CUDA kernel) break
them toward the LOWER index, which is what the synthetic frame generator (cofii2p_b200/frames.py) fixes too.
Pinned against the reference's own `knn()` run in the build container (tests/test_cpu.py::test_knn_oracle_vs_reference)
on integer-lattice clouds, where every distance is exact and the only freedom is the tie order.
Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this module.
"""
import numpy as np

DIRECT, EXPANDED = 0, 1
f32 = np.float32


def distances(src: np.ndarray, qry: np.ndarray, mode: int) -> np.ndarray:
    """[nq, ns] fp32 squared distances, each operation rounded to fp32 in the reference's order (numpy never fuses)."""
    s, q = src.astype(f32), qry.astype(f32)
    if mode == DIRECT:
        dx = q[:, None, 0] - s[None, :, 0]
        dy = q[:, None, 1] - s[None, :, 1]
        dz = q[:, None, 2] - s[None, :, 2]
        return (dx * dx + dy * dy) + dz * dz
    dot = (q[:, None, 0] * s[None, :, 0] + q[:, None, 1] * s[None, :, 1]) + q[:, None, 2] * s[None, :, 2]
    qq = (q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1]) + q[:, 2] * q[:, 2]
    ss = (s[:, 0] * s[:, 0] + s[:, 1] * s[:, 1]) + s[:, 2] * s[:, 2]
    d = f32(-2.0) * dot            # preprocess_data.py:121
    d = d + qq[:, None]            # :125
    d = d + ss[None, :]            # :126
    return np.maximum(d, f32(1e-12))  # :128


def knn_table(src: np.ndarray, qry: np.ndarray, k: int = 128, mode: int = DIRECT, chunk: int = 1024) -> np.ndarray:
    """[nq, k] int64, rows ascending in (distance, index); when ns < k the tail holds ns (shadow index)."""
    ns, nq = src.shape[0], qry.shape[0]
    out = np.full((nq, k), ns, dtype=np.int64)
    kk = min(k, ns)
    for a in range(0, nq, chunk):
        d = distances(src, qry[a:a + chunk], mode)
        order = np.argsort(d, axis=1, kind="stable")  # stable: equal distances keep ascending index
        out[a:a + chunk, :kk] = order[:, :kk]
    return out


def half_sample_pyramid(points: np.ndarray, num_stages: int, rng=np.random):
    """points [3, N] -> list of [N_i, 3] fp32 (preprocess_data.py:52-68): stage i>0 draws n//2 indices WITH replacement
    from the previous stage through the global numpy RNG."""
    levels = []
    for i in range(num_stages):
        if i > 0:
            idx = rng.choice(np.arange(points.shape[1]), size=points.shape[1] // 2)
            points = points[:, idx]
        levels.append(np.ascontiguousarray(points.T.astype(f32)))
    return levels


def pyramid_tables(levels, k: int = 128, mode: int = DIRECT):
    """dict(neighbors, subsampling, upsampling) of int64 tables (preprocess_data.py:75-99)."""
    L = len(levels)
    return {
        "neighbors": [knn_table(levels[i], levels[i], k, mode) for i in range(L)],
        "subsampling": [knn_table(levels[i], levels[i + 1], k, mode) for i in range(L - 1)],
        "upsampling": [knn_table(levels[i + 1], levels[i], k, mode) for i in range(L - 1)],
    }
