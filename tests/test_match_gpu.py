"""The matching stage on the product path (reference model/network.py:145-187, evaluation/eval_all.py:96-105):
exact tensor-core similarity + arg-min (tcgen05 fp16 candidate pass + exact fp32 re-rank) must be BIT-IDENTICAL to the
fp32 SIMT engine (itself pinned to the reference's arithmetic by test_ops_gpu.py::test_matching_exact and the goldens);
the fixed-shape batched test-mode tail and the graph-captured test-mode engine must reproduce the eager forward."""
import pytest
import torch
import torch.nn.functional as F

from conftest import get_frame, rel_err

pytestmark = pytest.mark.gpu


def _feats(rows, c, g, scale=1.0):
    return (F.normalize(torch.randn((rows, c), generator=g), dim=1) * scale).cuda()


@pytest.mark.parametrize("npt,npx,c,frames", [(1280, 1280, 128, 8), (1280, 1280, 128, 1), (1000, 3000, 64, 2), (300, 130, 64, 1),
                                              (37, 90, 128, 3), (512, 20480, 64, 1), (2560, 4000, 128, 1)])
def test_sim_argmin_exact_is_bit_identical(npt, npx, c, frames):
    from cofii2p_b200 import ops
    g = torch.Generator().manual_seed(npt + npx + c)
    pt, px = _feats(frames * npt, c, g), _feats(frames * npx, c, g)
    stats = torch.zeros(2, dtype=torch.int32, device="cuda")
    bi, bv = ops.sim_argmin(pt, px, frames, engine=ops.ENGINE_FP32)
    ti, tv = ops.sim_argmin(pt, px, frames, stats=stats)              # default engine: exact tensor-core path
    assert torch.equal(bi, ti) and torch.equal(bv, tv)
    cand, scans = stats.tolist()
    assert scans == 0 and frames * npt <= cand < 4 * frames * npt, (cand, scans)   # a handful of candidates per row
    # producer-supplied fp16 copies (the engine path: l2norm_rows_f16) give the same answer
    ptn, pth = ops.l2norm_rows_f16(pt)
    pxn, pxh = ops.l2norm_rows_f16(px)
    ui, uv = ops.sim_argmin(ptn, pxn, frames, pt_h=pth, px_h=pxh)
    ri, rv = ops.sim_argmin(ptn, pxn, frames, engine=ops.ENGINE_FP32)
    assert torch.equal(ui, ri) and torch.equal(uv, rv)


def test_sim_argmin_exact_near_ties_and_exact_ties():
    """Smooth image features: many pixels within 1e-4..1e-3 of the best (fp16 cannot order them) and exact duplicates
    (lowest index must win, as torch.argmin / the fp32 engine pick it)."""
    from cofii2p_b200 import ops
    g = torch.Generator().manual_seed(5)
    npt, npx, c = 640, 1280, 128
    base = F.normalize(torch.randn((64, c), generator=g), dim=1)
    px = base[torch.randint(0, 64, (npx,), generator=g)]
    px = F.normalize(px + 2e-4 * torch.randn((npx, c), generator=g), dim=1)   # clusters of ~20 near-identical pixels
    px[700:720] = px[100:120]                                                  # exact duplicates at higher indices
    pt = F.normalize(base[torch.randint(0, 64, (npt,), generator=g)] + 0.05 * torch.randn((npt, c), generator=g), dim=1)
    pt, px = pt.cuda().contiguous(), px.cuda().contiguous()
    stats = torch.zeros(2, dtype=torch.int32, device="cuda")
    bi, bv = ops.sim_argmin(pt, px, 1, engine=ops.ENGINE_FP32)
    ti, tv = ops.sim_argmin(pt, px, 1, stats=stats)
    assert torch.equal(bi, ti) and torch.equal(bv, tv)
    assert int(stats[0]) > 4 * npt          # the candidate lists really were exercised
    assert not bool(((ti >= 700) & (ti < 720)).any())


def test_sim_argmin_exact_overflow_falls_back_to_full_scan():
    """More than 16 pixels inside the margin of one pixel-range split: the row is scanned exactly in full."""
    from cofii2p_b200 import ops
    g = torch.Generator().manual_seed(9)
    npt, npx, c = 300, 2048, 64
    one = F.normalize(torch.randn((1, c), generator=g), dim=1)
    px = F.normalize(one + 1e-4 * torch.randn((npx, c), generator=g), dim=1).cuda()   # every pixel within the margin
    pt = _feats(npt, c, g)
    stats = torch.zeros(2, dtype=torch.int32, device="cuda")
    bi, bv = ops.sim_argmin(pt, px, 1, engine=ops.ENGINE_FP32)
    ti, tv = ops.sim_argmin(pt, px, 1, stats=stats)
    assert torch.equal(bi, ti) and torch.equal(bv, tv)
    assert int(stats[1]) == npt


@pytest.mark.parametrize("scale_pt,scale_px", [(3.7, 1.0), (0.01, 40.0), (1e-3, 1e-3)])
def test_sim_argmin_exact_unnormalised_rows(scale_pt, scale_px):
    """Rows that are not unit-norm: the margin scales with the measured norm bound (cofi_cast_f16_bound)."""
    from cofii2p_b200 import ops
    g = torch.Generator().manual_seed(13)
    pt, px = _feats(700, 128, g, scale_pt), _feats(1500, 128, g, scale_px)
    pt = pt * (0.5 + torch.rand((700, 1), generator=g).cuda())
    bi, bv = ops.sim_argmin(pt, px, 1, engine=ops.ENGINE_FP32)
    ti, tv = ops.sim_argmin(pt, px, 1)
    assert torch.equal(bi, ti) and torch.equal(bv, tv)


def test_sim_argmin_exact_vs_reference_arithmetic():
    """Directly against the reference's expression (network.py:174-179) evaluated by torch on the CPU."""
    from cofii2p_b200 import ops
    g = torch.Generator().manual_seed(21)
    img = F.normalize(torch.randn((1, 128, 20, 64), generator=g), dim=1)
    pc = F.normalize(torch.randn((128, 1280), generator=g), dim=0)
    dist = 1 - torch.sum(img.flatten(2)[0].unsqueeze(-1) * pc.unsqueeze(-2), dim=0)
    best, val = ops.sim_argmin(pc.t().contiguous().cuda(), img.flatten(2)[0].t().contiguous().cuda(), 1)
    assert torch.equal(best.cpu(), torch.argmin(dist, dim=0))
    assert torch.equal(val.cpu(), dist.min(0).values)


def _select_ref(score, best, H, W, thresholds, min_count, xy_scale):
    x, y = best % W, best // W
    m = (x >= 2) & (x <= 62) & (y <= 18) & (y >= 2)
    chosen = len(thresholds) - 1
    for t, thr in enumerate(thresholds.tolist()):
        if int(((score >= thr) & m).sum()) >= min_count:
            chosen = t
            break
    keep = (score >= thresholds[chosen]) & m
    idx = torch.nonzero(keep).squeeze(1)
    return idx, torch.stack([x[idx].float() * xy_scale, y[idx].float() * xy_scale]), chosen


@pytest.mark.parametrize("npt,frames,lo,hi", [(1280, 8, 0.0, 1.0), (1280, 2, 0.4, 0.6), (3000, 1, 0.0, 0.05), (100, 3, 0.95, 1.0)])
def test_select_matches_parallel_compaction(npt, frames, lo, hi):
    from cofii2p_b200 import ops
    from cofii2p_b200.model.network import _thresholds
    g = torch.Generator().manual_seed(npt + frames)
    score = (lo + (hi - lo) * torch.rand((frames * npt,), generator=g)).cuda()
    best = torch.randint(0, 20 * 64, (frames * npt,), generator=g).cuda()
    thr = _thresholds("cuda")
    cnt, oidx, oxy = ops.select_matches(score, best, frames, 20, 64, thr, 4, xy_scale=4.0)
    for f in range(frames):
        ridx, rxy, chosen = _select_ref(score[f * npt:(f + 1) * npt], best[f * npt:(f + 1) * npt], 20, 64, thr, 4, 4.0)
        n = int(cnt[f, 0])
        assert n == ridx.numel() and int(cnt[f, 1]) == chosen
        assert torch.equal(oidx[f, :n], ridx) and torch.equal(oxy[f, :, :n], rxy)
        assert bool((oidx[f, n:] == 0).all()) and bool((oxy[f, :, n:] == 8.0).all())     # valid padding


def test_batched_tail_kernels_match_per_frame():
    from cofii2p_b200 import ops
    g = torch.Generator().manual_seed(3)
    B, n, M = 3, 200, 5000
    nodes = ((torch.rand((B * M, 3), generator=g) - 0.5) * 100).cuda()
    pts = ((torch.rand((B * n, 3), generator=g) - 0.5) * 100).cuda()
    got = ops.nn_argmin_batched(pts, nodes, B)
    for b in range(B):
        assert torch.equal(got[b * n:(b + 1) * n], ops.nn_argmin(pts[b * n:(b + 1) * n], nodes[b * M:(b + 1) * M]))
    fmap = torch.randn((B, 80, 256, 64), generator=g).cuda()
    ctr = torch.stack([torch.randint(2, 254, (B, n), generator=g), torch.randint(2, 78, (B, n), generator=g)], 1).float().cuda()
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    got = ops.extract_patch_batched(fmap, ctr, err)
    for b in range(B):
        assert torch.equal(got[b], ops.extract_patch(fmap, b, ctr[b].contiguous(), err))
    assert int(err) == 0


@pytest.mark.parametrize("engine", ["fp32", "tf32"])
def test_engine_test_mode_matches_eager_forward(cuda_model, engine):
    """mode='test' inside the captured graph (fixed shapes, device-side counts) == the per-frame eager forward(test):
    descriptors/scores equal, correspondences identical; also after a second batch is uploaded."""
    from cofii2p_b200 import evaluate as ev, ops
    from cofii2p_b200.engine import InferenceEngine
    from cofii2p_b200.frames import frame_to, stack_frames
    ops.set_engine(engine)
    try:
        fa = [get_frame(s, 4096) for s in (0, 1)]
        fb = [get_frame(s, 4096) for s in (1, 2)]
        eng = InferenceEngine(cuda_model, stack_frames(fa), mode="test", use_graph=True)
        assert eng.graph is not None
        for frames in (fa, fb):
            eng.upload(eng.host_buffers(stack_frames(frames)))
            eng.run()
            eng.download()
            outs = eng.results()
            corr = eng.correspondences()
            for f, o, c in zip(frames, outs, corr):
                fd = frame_to(f, "cuda")
                with torch.no_grad():
                    single = cuda_model(fd["pc_data_dict"], fd["img"], fd["fine_center_kpt_coors"], fd["fine_xy"],
                                        fd["fine_pc_inline_index"], "test")
                for a, b in zip(o, single):
                    assert tuple(a.shape) == tuple(b.shape) and rel_err(a, b) < 1e-6
                assert torch.equal(o[6], single[6]) and torch.equal(o[7], single[7])
                ip, op_, idx = ev.correspondences(single)
                assert torch.equal(c[0].cpu(), torch.from_numpy(ip)) and torch.equal(c[1].cpu(), torch.from_numpy(op_))
                assert torch.equal(c[2], idx)
    finally:
        ops.set_engine("fp32")


def test_engine_test_mode_reproduces_reference_golden(cuda_model):
    """fp32 engine, graph-captured test mode at the full KITTI size against the REAL reference's frozen outputs."""
    import os
    import numpy as np
    from cofii2p_b200 import ops
    from cofii2p_b200.engine import InferenceEngine
    from cofii2p_b200.frames import stack_frames
    ops.set_engine("fp32")
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frame_s0_n20480.npz"))
    eng = InferenceEngine(cuda_model, stack_frames([get_frame(0, 20480)]), mode="test", use_graph=True)
    eng.run()
    out = eng.results()[0]
    assert torch.equal(out[6].cpu(), torch.from_numpy(z["test/fine_center_xy"]))
    assert torch.equal(out[7].cpu(), torch.from_numpy(z["test/coarse_pc_points"]))
    for i, nm in ((4, "fine_img_feature_patch"), (5, "fine_pc_inline_feature")):
        g = torch.from_numpy(z["test/" + nm])
        assert tuple(out[i].shape) == tuple(g.shape) and rel_err(out[i], g) < 1e-3


# ---------------------------------------------------------------------------------------------- pose step (row f2)
def test_pnp_ransac_kernel_vs_oracle_and_opencv():
    """csrc/pnp.cu: (1) against its numpy restatement (oracle/pnp.py) on the same counter-based samples -- same winning
    hypothesis, inlier count, inlier mask, pose to 1e-9; (2) against the reference's solver call cv2.solvePnPRansac
    (evaluation/eval_all.py:107) on synthetic correspondences with a known pose: identical inlier set, hence the identical
    refined pose (RTE / RRE difference far below the north star's 1e-3); batched frames with ragged counts."""
    import cv2
    import numpy as np
    from cofii2p_b200 import evaluate as ev, ops
    from oracle import pnp
    sys_case = __import__("test_cpu")._pnp_case
    cases = [sys_case(s, n) for s, n in ((1, 300), (2, 300), (3, 120))]
    n_max = 320
    B = len(cases)
    ip = torch.zeros((B, n_max, 2))
    op = torch.zeros((B, n_max, 3))
    cnt = torch.zeros((B, 2), dtype=torch.int32)
    for b, (K, img, obj, R, t, truth) in enumerate(cases):
        n = img.shape[0]
        ip[b, :n], op[b, :n], cnt[b, 0] = torch.from_numpy(img), torch.from_numpy(obj), n
    K = cases[0][0]
    cam = torch.tensor([[K[0, 0], K[1, 1], K[0, 2], K[1, 2]]] * B, dtype=torch.float32).cuda()
    iters = 600
    o_cnt, o_hyp, o_pose, o_inl = ops.pnp_ransac(ip.cuda(), op.cuda(), cam, cnt.cuda(), iters, 8.0, seed=7)
    for b, (K, img, obj, R, t, truth) in enumerate(cases):
        n = img.shape[0]
        ref = pnp.ransac(K, img, obj, iterations=iters, threshold=8.0, seed=7)
        assert int(o_cnt[b]) == ref["count"] and int(o_hyp[b]) == ref["hypothesis"]
        assert np.array_equal(o_inl[b, :n].cpu().numpy().astype(bool), ref["inliers"]) and int(o_inl[b, n:].sum()) == 0
        pose = o_pose[b].cpu().numpy()
        assert np.abs(pose[:9].reshape(3, 3) - ref["R"]).max() < 1e-9 and np.abs(pose[9:] - ref["t"]).max() < 1e-8
    # the full pose step (10000 hypotheses) against OpenCV's
    res = ev.solve_pose_batch(K, ip.cuda(), op.cuda(), cnt.cuda(), iterations=10000, threshold=8.0, seed=0)
    for b, (K, img, obj, R, t, truth) in enumerate(cases):
        ok_cv, T_cv, inl_cv = ev.solve_pose(K, img, obj)
        ok, T, inl = res[b]
        assert ok and ok_cv and set(inl_cv[:, 0].tolist()) == set(inl.tolist())
        T_gt = np.eye(4)
        T_gt[:3, :3], T_gt[:3, 3] = R, t
        rte, rre = ev.pose_error(T, T_gt)
        rte_cv, rre_cv = ev.pose_error(T_cv, T_gt)
        assert abs(rte - rte_cv) < 1e-6 and abs(rre - rre_cv) < 1e-6 and rte < 0.05 and rre < 0.5
    ok1, T1, inl1 = ev.solve_pose_gpu(K, cases[0][1], cases[0][2])
    assert ok1 and np.allclose(T1, res[0][1], atol=1e-9) and inl1.shape[1] == 1


def test_engine_pose_step_on_device(cuda_model):
    """Graph-captured test-mode engine -> padded device correspondences -> batched RANSAC: one launch for all frames; the
    padded rows never count.  (Random-weight descriptors give no geometric consensus, so neither the poses nor the chance
    inlier counts of the two solvers are comparable on these frames -- the synthetic case above is the parity test.)"""
    import numpy as np
    from cofii2p_b200 import evaluate as ev, ops
    from cofii2p_b200.engine import InferenceEngine
    from cofii2p_b200.frames import stack_frames
    ops.set_engine("fp32")
    frames = [get_frame(s, 4096) for s in (0, 1)]
    eng = InferenceEngine(cuda_model, stack_frames(frames), mode="test", use_graph=True)
    eng.run()
    ip, op_, cnt = eng.correspondences_padded()
    K = frames[0]["K_half"].numpy()
    res = ev.solve_pose_batch(K, ip, op_, cnt, iterations=10000, threshold=8.0, seed=0)
    corr = eng.correspondences()
    for b, (ok, T, inl) in enumerate(res):
        n = int(cnt[b, 0])
        assert torch.equal(ip[b, :n], corr[b][0]) and torch.equal(op_[b, :n], corr[b][1])
        assert len(inl) >= 3 and int(inl.max()) < n and T.shape == (4, 4) and np.isfinite(T).all()
