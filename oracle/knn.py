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


def philox4x32_10_word0(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al., SC'11), first output word, vectorised over numpy uint32 arrays."""
    u32, u64 = np.uint32, np.uint64
    c0, c1, c2, c3 = (np.asarray(c, dtype=u32) for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = u32(k0), u32(k1)
    for _ in range(10):
        p0 = u64(0xD2511F53) * c0.astype(u64)
        p1 = u64(0xCD9E8D57) * c2.astype(u64)
        n0 = (p1 >> u64(32)).astype(u32) ^ c1 ^ k0
        n1 = p1.astype(u32)
        n2 = (p0 >> u64(32)).astype(u32) ^ c3 ^ k1
        n3 = p0.astype(u32)
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = u32((int(k0) + 0x9E3779B9) & 0xFFFFFFFF)
        k1 = u32((int(k1) + 0xBB67AE85) & 0xFFFFFFFF)
    return c0


def half_sample_pyramid_philox(points0: np.ndarray, num_stages: int, seed: int, frame: int = 0):
    """The device sampler's definition (csrc/sample.cu), restated: points0 [N, 3] -> (levels [N_i, 3], level-0 indices).
    Same distribution as preprocess_data.py:58 (n // 2 indices WITH replacement per stage), drawn from a counter-based
    generator instead of numpy's global stream: index_j = (Philox(j, frame, level, 0; seed)[0] * n_prev) >> 32."""
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    levels, index = [np.ascontiguousarray(points0, dtype=f32)], [np.arange(points0.shape[0], dtype=np.int64)]
    for l in range(1, num_stages):
        n_prev = levels[-1].shape[0]
        j = np.arange(n_prev // 2, dtype=np.uint32)
        u = philox4x32_10_word0(j, np.uint32(frame), np.uint32(l), np.uint32(0), k0, k1)
        pick = ((u.astype(np.uint64) * np.uint64(n_prev)) >> np.uint64(32)).astype(np.int64)
        levels.append(levels[-1][pick])
        index.append(index[-1][pick])
    return levels, index


def pyramid_tables(levels, k: int = 128, mode: int = DIRECT):
    """dict(neighbors, subsampling, upsampling) of int64 tables (preprocess_data.py:75-99)."""
    L = len(levels)
    return {
        "neighbors": [knn_table(levels[i], levels[i], k, mode) for i in range(L)],
        "subsampling": [knn_table(levels[i], levels[i + 1], k, mode) for i in range(L - 1)],
        "upsampling": [knn_table(levels[i + 1], levels[i], k, mode) for i in range(L - 1)],
    }
