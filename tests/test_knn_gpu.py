"""GPU parity of the pyramid / KNN table builder (csrc/knn.cu), through the C ABI.

Checker = oracle/knn.py (numpy restatement of reference model/kpconv/preprocess_data.py) at sizes it finishes in
seconds, the frozen outputs of the reference's own `knn()` / stack-mode driver (tests/golden/knn_ref.npz), and at the
full 20480-point size a same-arithmetic torch brute force plus size-independent properties (sorted rows, self first,
nothing outside the row is closer than its last entry).  Index work: bit-exact."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def _clouds(kind, rng, n):
    if kind == "lattice":          # exact distances, many ties
        p = rng.integers(-50, 50, (n, 3)).astype(np.float32)
        p[:, 1] = np.round(p[:, 1] / 8)
    elif kind == "float":
        p = (rng.normal(size=(n, 3)) * np.array([20, 2, 20])).astype(np.float32)
    elif kind == "far":            # 80 m from the origin: the expanded form's cancellation noise is ~1e-3 m^2
        p = (rng.normal(size=(n, 3)) * np.array([6, 1, 6]) + np.array([80, 1, -75])).astype(np.float32)
    else:                          # duplicates, as the reference's sampling WITH replacement produces
        p = (rng.normal(size=(n // 2 + 1, 3)) * 10).astype(np.float32)
        p = p[rng.integers(0, p.shape[0], n)]
    return p


@pytest.mark.parametrize("kind", ["lattice", "float", "far", "dups"])
@pytest.mark.parametrize("mode", [0, 1])
def test_knn_table_vs_oracle(kind, mode):
    from cofii2p_b200 import ops
    from oracle import knn as ok
    rng = np.random.default_rng(["lattice", "float", "far", "dups"].index(kind) * 2 + mode)
    for ns, nq, k in ((1500, 700, 128), (333, 1000, 128), (100, 37, 128), (1, 5, 128), (4096, 300, 16), (257, 257, 1)):
        src, qry = _clouds(kind, rng, ns), _clouds(kind, rng, nq)
        if nq <= ns and kind != "far":
            qry[: nq // 2] = src[rng.permutation(ns)[: nq // 2]]     # half the queries are members of the source
        want = ok.knn_table(src, qry, k, mode)
        for flags in (0, ops.KNN_NOCULL, ops.KNN_NOFAST):   # fast (threshold selection), brute force, iterative selection
            got = ops.knn_table(_cuda(src), _cuda(qry), 1, k, mode | flags).cpu().numpy()
            assert np.array_equal(got, want), (kind, mode, ns, nq, k, flags, int((got != want).sum()))


def test_knn_table_self_and_frames():
    """src is qry (same-level job), several stacked frames with frame-local indices."""
    from cofii2p_b200 import ops
    from oracle import knn as ok
    rng = np.random.default_rng(2)
    B, n = 3, 900
    clouds = [_clouds("lattice" if b == 0 else "float", rng, n) for b in range(B)]
    x = _cuda(np.concatenate(clouds, 0))
    got = ops.knn_table(x, x, B, 128, 0).cpu().numpy()
    for b in range(B):
        assert np.array_equal(got[b * n:(b + 1) * n], ok.knn_table(clouds[b], clouds[b], 128, 0)), b


@pytest.mark.parametrize("mode", [0, 1])
def test_knn_pyramid_vs_oracle(mode):
    from cofii2p_b200 import ops
    from oracle import knn as ok
    rng = np.random.default_rng(4)
    B, n0, L = 2, 2000, 4
    per_frame = []
    for b in range(B):
        np.random.seed(20 + b)
        per_frame.append(ok.half_sample_pyramid(_clouds("float" if b else "lattice", rng, n0).T, L))
    levels = [_cuda(np.concatenate([per_frame[b][l] for b in range(B)], 0)) for l in range(L)]
    got = ops.knn_pyramid(levels, frames=B, k=128, mode=mode)
    for b in range(B):
        want = ok.pyramid_tables(per_frame[b], 128, mode)
        for name in ("neighbors", "subsampling", "upsampling"):
            for l, t in enumerate(want[name]):
                g = got[name][l].cpu().numpy()
                rows = t.shape[0]
                assert np.array_equal(g[b * rows:(b + 1) * rows], t), (name, l, b)
    # a subset of the tables only
    part = ops.knn_pyramid(levels, frames=B, k=128, mode=mode, want=("upsampling",))
    assert part["neighbors"] == [] and torch.equal(part["upsampling"][1], got["upsampling"][1])


def test_knn_vs_reference_golden():
    """Outputs of the reference's own knn() / precompute_point_cloud_stack_mode (frozen by oracle/make_knn_golden.py)."""
    from cofii2p_b200 import ops
    from oracle import knn as ok
    from oracle import make_knn_golden as mk
    import model.kpconv.preprocess_data as pp
    z = np.load(os.path.join(GOLD, "knn_ref.npz"))
    src, qry = mk.knn_case()
    got = pp.knn(torch.from_numpy(src), torch.from_numpy(qry), 128).numpy()
    ref = z["knn_idx"].astype(np.int64)
    d = ok.distances(src, qry, ok.EXPANDED)
    dr, dg = np.take_along_axis(d, ref, 1), np.take_along_axis(d, got, 1)
    assert np.array_equal(dr, dg)
    strict = dg < dg[:, -1:]
    for r in range(ref.shape[0]):
        assert set(ref[r][strict[r]]) == set(got[r][strict[r]])
    # the drop-in driver, seeded like the golden run: identical pyramid and identical tables
    np.random.seed(7)
    out = pp.precompute_point_cloud_stack_mode(mk.driver_case(), None, None, lengths=2048, num_stages=5)
    assert out["lengths"] == list(z["lengths"])
    for i, p in enumerate(out["points"]):
        assert p.is_cuda and np.array_equal(p.cpu().numpy(), z[f"points{i}"])
    for name in ("neighbors", "subsampling", "upsampling"):
        for i, t in enumerate(out[name]):
            assert t.dtype == torch.int64 and np.array_equal(t.cpu().numpy(), z[f"{name}{i}"].astype(np.int64)), (name, i)


def _torch_direct_rows(src, qry, rows):
    """same arithmetic as COFI_KNN_DIRECT with torch elementwise ops (each rounded to fp32), sorted by (d, index)."""
    q = qry[rows]
    dx, dy, dz = q[:, None, 0] - src[None, :, 0], q[:, None, 1] - src[None, :, 1], q[:, None, 2] - src[None, :, 2]
    d = (dx * dx + dy * dy) + dz * dz
    dv, di = torch.sort(d, dim=1, stable=True)
    return dv, di


def test_knn_full_size_frame():
    """20480-point synthetic KITTI frame: all 13 tables; checked against the frame generator's exact-integer tables where
    the geometry is still on the lattice, and against a same-arithmetic torch brute force on the posed (float) clouds."""
    from cofii2p_b200 import ops
    from cofii2p_b200.frames import make_frame
    from conftest import FRAME_CACHE
    f = make_frame(3, num_pc=20480, cache_dir=FRAME_CACHE, device="cuda")
    d = f["pc_data_dict"]
    levels = [p.cuda() for p in d["points"]]
    got = ops.knn_pyramid(levels, frames=1, k=128, mode=0)
    nocull = ops.knn_pyramid(levels, frames=1, k=128, mode=ops.KNN_NOCULL)
    nofast = ops.knn_pyramid(levels, frames=1, k=128, mode=ops.KNN_NOFAST)
    g = torch.Generator().manual_seed(0)
    for name, pairs in (("neighbors", [(l, l) for l in range(5)]), ("subsampling", [(l, l + 1) for l in range(4)]),
                        ("upsampling", [(l + 1, l) for l in range(4)])):
        for i, (s, q) in enumerate(pairs):
            t = got[name][i]
            assert torch.equal(t, nocull[name][i]), (name, i)
            assert torch.equal(t, nofast[name][i]), (name, i)
            rows = torch.randperm(levels[q].shape[0], generator=g)[:256].cuda()
            dv, di = _torch_direct_rows(levels[s], levels[q], rows)
            assert torch.equal(t[rows], di[:, :128]), (name, i)
            # the posed cloud is a rigid motion of the lattice: the generator's exact tables differ only inside groups of (near-)equal distance, which the lattice is full of
            assert (t.cpu() == d[name][i]).float().mean() > 0.8, (name, i)
    assert torch.equal(got["neighbors"][0][:, 0].cpu(), torch.arange(20480))  # self first (no duplicate points)


def test_knn_integer_cloud_equals_frame_generator():
    """On integer-valued coordinates the kernel reproduces frames._knn_table bit for bit, so a forward pass fed with
    device-built tables is the forward pass of the committed goldens."""
    from cofii2p_b200 import ops
    from cofii2p_b200.frames import _knn_table
    rng = np.random.default_rng(9)
    lat = np.unique(rng.integers(-120, 120, (9000, 3)), axis=0).astype(np.int32)
    lat[:, 1] //= 10
    lat = np.unique(lat, axis=0)
    rng.shuffle(lat, axis=0)
    sub = lat[: lat.shape[0] // 2]
    tl, ts = torch.from_numpy(lat).cuda(), torch.from_numpy(sub).cuda()
    got = ops.knn_pyramid([tl.float(), ts.float()], frames=1, k=128, mode=0)
    assert torch.equal(got["neighbors"][0], _knn_table(tl, tl, 128))
    assert torch.equal(got["subsampling"][0], _knn_table(tl, ts, 128))
    assert torch.equal(got["upsampling"][0], _knn_table(ts, tl, 128))


def test_knn_bad_arguments():
    from cofii2p_b200 import ops
    x = torch.zeros((64, 3), device="cuda")
    with pytest.raises(RuntimeError):
        ops.knn_table(x, x, 1, 129, 0)
    with pytest.raises(RuntimeError):
        ops.knn_table(x, x, 1, 16, 7)
    with pytest.raises(RuntimeError):
        ops.knn_table(x.cpu(), x, 1, 16, 0)


def test_engine_builds_tables_on_device():
    """InferenceEngine(tables='device'): the captured graph starts from the point pyramid alone and equals the engine that
    is handed the same tables."""
    from cofii2p_b200 import ops
    from cofii2p_b200.engine import InferenceEngine
    from cofii2p_b200.frames import make_frame, stack_frames
    from cofii2p_b200.model.network import CoFiI2P
    from cofii2p_b200.options import Options_KITTI
    from cofii2p_b200.weights import seeded_state_dict
    from conftest import FRAME_CACHE
    m = CoFiI2P(Options_KITTI())
    m.load_state_dict(seeded_state_dict(m, 0), strict=True)
    m = m.cuda().eval()
    ops.set_engine("fp32")
    frames = [make_frame(s, num_pc=4096, cache_dir=FRAME_CACHE, device="cuda") for s in (0, 1)]
    batch = stack_frames(frames)
    pts = [p.cuda() for p in batch["pc_data_dict"]["points"]]
    tabs = ops.knn_pyramid(pts, frames=2, k=128)
    for name in ("neighbors", "subsampling", "upsampling"):
        batch["pc_data_dict"][name] = tabs[name]
    a = InferenceEngine(m, batch, tables="host")
    b = InferenceEngine(m, batch, tables="device")
    assert b.launches_per_step == a.launches_per_step + 4   # two sorts, the query launch, the sub-sampling (row copy) launch
    for e in (a, b):
        e.run()
    for ra, rb in zip(a.results(), b.results()):
        for x, y in zip(ra[:6], rb[:6]):
            assert torch.equal(x, y)
    hb = b.host_buffers(batch)
    assert hb["neighbors"] == [] and b.upload(hb) < 2 * (4096 * 2 * 3 * 4 + 4096 * 16 + 3 * 160 * 512 * 4) + 4096
    # another batch: the table graph (its own stream) must rebuild the tables from the uploaded pyramid every step, also
    # inside the pipelined engine where it runs beside the previous batch's forward
    from cofii2p_b200.engine import PipelinedEngine
    batch2 = stack_frames([frames[1], frames[0]])
    tabs2 = ops.knn_pyramid([p.cuda() for p in batch2["pc_data_dict"]["points"]], frames=2, k=128)
    for name in ("neighbors", "subsampling", "upsampling"):
        batch2["pc_data_dict"][name] = [t.cpu() for t in tabs2[name]]     # host_buffers() pins host tensors
    batch = stack_frames(frames)                                          # (the first batch again, host tensors throughout)
    for name in ("neighbors", "subsampling", "upsampling"):
        batch["pc_data_dict"][name] = [t.cpu() for t in tabs[name]]
    a.upload(a.host_buffers(batch2))
    b.upload(b.host_buffers(batch2))
    for e in (a, b):
        e.run()
    ref2 = a.results()
    for ra, rb in zip(ref2, b.results()):
        for x, y in zip(ra[:6], rb[:6]):
            assert torch.equal(x, y)
    ref1 = InferenceEngine(m, batch, tables="host")
    ref1.run()
    ref1 = ref1.results()
    pipe = PipelinedEngine(m, batch, depth=2, tables="device")
    hosts = [pipe.engines[0].host_buffers(bt) for bt in (batch, batch2)]
    for i in range(5):
        pipe.step(hosts[i % 2])
        for ra, rb in zip(ref1 if i % 2 == 0 else ref2, pipe.last_results()):
            for x, y in zip(ra[:6], rb[:6]):
                assert torch.equal(x, y)


def test_precompute_point_cloud_cuda_surface():
    """The reference's second builder (model/kpconv/preprocess_data.py:145-203): half-sampling without replacement and the
    expanded-form `knn()` ranking.  open3d's private RNG cannot be reproduced, so the check is on the returned pyramid:
    every level is a subset of the previous one without duplicates, and the 13 tables are the oracle's for those levels."""
    import model.kpconv.preprocess_data as pp
    from oracle import knn as ok
    rng = np.random.default_rng(21)
    pts = (rng.normal(size=(3, 2048)) * np.array([[15.0], [1.5], [15.0]])).astype(np.float32)
    np.random.seed(3)
    out = pp.precompute_point_cloud_cuda(pts, None, None, lengths=2048, num_stages=5, device="cpu")
    assert out["lengths"] == [2048, 1024, 512, 256, 128]
    levels = [p.numpy() for p in out["points"]]
    assert not out["points"][0].is_cuda and np.array_equal(levels[0], pts.T)
    for a, b in zip(levels[:-1], levels[1:]):
        assert b.shape[0] == a.shape[0] // 2 and np.unique(b, axis=0).shape[0] == b.shape[0]
        assert set(map(tuple, b.tolist())) <= set(map(tuple, a.tolist()))
    want = ok.pyramid_tables(levels, 128, ok.EXPANDED)
    for name in ("neighbors", "subsampling", "upsampling"):
        for i, t in enumerate(want[name]):
            assert np.array_equal(out[name][i].numpy(), t), (name, i)


@pytest.mark.parametrize("n0,frames,levels,seed", [(20480, 2, 5, 0), (4096, 3, 5, 12345678901234), (1000, 1, 4, 7)])
def test_device_half_sampling_matches_the_oracle(n0, frames, levels, seed):
    """cofi_half_sample_pyramid (reference preprocess_data.py:52-68 semantics: n//2 draws WITH replacement per stage; the
    draw itself is the counter-based Philox definition) against oracle/knn.py::half_sample_pyramid_philox, bit for bit,
    including the level-0 index of every sampled row; and the sampled pyramid feeds the table builder."""
    import numpy as np
    from cofii2p_b200 import ops
    from oracle import knn as ok
    g = torch.Generator().manual_seed(n0)
    pts = (torch.rand((frames * n0, 3), generator=g) - 0.5) * 100
    levels_d, index_d = ops.half_sample_pyramid(pts.cuda(), frames, levels, seed, want_index=True)
    assert levels_d[0].data_ptr() != 0 and len(levels_d) == levels
    for f in range(frames):
        ref_l, ref_i = ok.half_sample_pyramid_philox(pts[f * n0:(f + 1) * n0].numpy(), levels, seed, frame=f)
        for l in range(1, levels):
            nl = n0 >> l
            assert tuple(levels_d[l].shape) == (frames * nl, 3)
            assert np.array_equal(levels_d[l][f * nl:(f + 1) * nl].cpu().numpy(), ref_l[l])
            assert np.array_equal(index_d[l][f * nl:(f + 1) * nl].cpu().numpy(), ref_i[l])
    # the draw is with replacement: duplicates exist -- 2 (1 - e^-0.5) = 79 % of the n/2 draws are distinct
    hit = torch.unique(index_d[1][:n0 // 2]).numel() / (n0 // 2)
    assert 0.70 < hit < 0.87, hit
    tabs = ops.knn_pyramid(levels_d, frames=frames, k=16)
    assert tabs["neighbors"][1].shape == (frames * (n0 >> 1), 16)
