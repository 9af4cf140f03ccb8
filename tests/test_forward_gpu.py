"""End-to-end parity of `CoFiI2P.forward` on the CUDA path: (1) against the CPU oracle run on the same box at
4096 points, (2) against golden outputs of the real reference (tests/golden, produced by oracle/make_golden.py)
at 4096 and at the full 20480-point KITTI size.  Tolerance: 1e-3 relative (north star), indices exact."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, get_frame, rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-3
NAMES = ["img_feature_norm", "pc_feature_norm", "coarse_img_score", "coarse_pc_score", "fine_img_feature_patch",
         "fine_pc_inline_feature", "fine_center_xy", "coarse_pc_points"]


def _run(model, frame, mode, taps=None):
    from cofii2p_b200.frames import frame_to
    from cofii2p_b200 import ops
    ops.set_engine("fp32")
    f = frame_to(frame, "cuda")
    with torch.no_grad():
        return model(f["pc_data_dict"], f["img"], f["fine_center_kpt_coors"], f["fine_xy"],
                     f["fine_pc_inline_index"], mode, taps=taps)


def test_forward_vs_oracle_same_box(cuda_model, seeded_sd):
    from oracle import restate
    frame = get_frame(0, 4096)
    for mode in ("val", "test"):
        with torch.no_grad():
            ref = restate.forward(seeded_sd, frame["pc_data_dict"], frame["img"], frame["fine_center_kpt_coors"],
                                  frame["fine_xy"], frame["fine_pc_inline_index"], mode)
        got = _run(cuda_model, frame, mode)
        for nm, a, b in zip(NAMES, got, ref):
            if b is None:
                assert a is None
                continue
            assert tuple(a.shape) == tuple(b.shape), (nm, a.shape, b.shape)
            assert rel_err(a, b) < TOL, (mode, nm, rel_err(a, b))


@pytest.mark.parametrize("seed,num_pc", [(0, 4096), (1, 4096), (0, 20480)])
def test_forward_vs_reference_golden(cuda_model, seed, num_pc):
    z = np.load(os.path.join(GOLD, f"frame_s{seed}_n{num_pc}.npz"))
    frame = get_frame(seed, num_pc)
    taps = {}
    val = _run(cuda_model, frame, "val", taps)
    test = _run(cuda_model, frame, "test")
    for i, nm in enumerate(NAMES):
        if val[i] is not None:
            g = torch.from_numpy(z["val/" + nm])
            assert tuple(val[i].shape) == tuple(g.shape)
            assert rel_err(val[i], g) < TOL, ("val", nm, rel_err(val[i], g))
    for i, nm in enumerate(NAMES[4:], 4):
        g = torch.from_numpy(z["test/" + nm])
        assert tuple(test[i].shape) == tuple(g.shape), (nm, test[i].shape, g.shape)
        assert rel_err(test[i], g) < TOL, ("test", nm, rel_err(test[i], g))
    # the selected correspondences (pixel centres, super-points) must be identical
    assert torch.equal(test[6].cpu(), torch.from_numpy(z["test/fine_center_xy"]))
    assert torch.equal(test[7].cpu(), torch.from_numpy(z["test/coarse_pc_points"]))
    # strided samples of intermediates localise a failure
    tapmap = {"pc_encoder.encoder1_2": "encoder1_2", "pc_encoder.encoder2_3": "encoder2_3",
              "pc_encoder.encoder3_3": "encoder3_3", "pc_encoder.encoder4_3": "encoder4_3",
              "pc_encoder.encoder5_3": "encoder5_3", "pc_encoder.decoder4": "decoder4",
              "pc_encoder.decoder3": "decoder3", "pc_encoder.decoder2": "decoder2"}
    for gname, tname in tapmap.items():
        g = torch.from_numpy(z["tap/" + gname])
        t = taps[tname].reshape(-1)
        step = max(1, t.numel() // 4096)
        assert rel_err(t[::step][:4096], g) < TOL, (gname, rel_err(t[::step][:4096], g))


def test_forward_batch_matches_single(cuda_model):
    from cofii2p_b200.frames import frame_to, stack_frames
    from cofii2p_b200 import ops
    ops.set_engine("fp32")
    frames = [get_frame(s, 4096) for s in (0, 1)]
    batch = frame_to(stack_frames(frames), "cuda")
    with torch.no_grad():
        outs = cuda_model.forward_batch(batch, "val")
    for f, o in zip(frames, outs):
        single = _run(cuda_model, f, "val")
        for a, b in zip(o, single):
            if b is None:
                assert a is None
            else:
                assert rel_err(a, b) < 1e-6
    with torch.no_grad():
        outs_t = cuda_model.forward_batch(batch, "test")
    for f, o in zip(frames, outs_t):
        single = _run(cuda_model, f, "test")
        assert torch.equal(o[6], single[6]) and torch.equal(o[7], single[7])


def test_fine_match_and_eval_tail(cuda_model):
    """pixel<->point assembly of evaluation/eval_all.py:99-105 on the test-mode outputs vs the oracle."""
    from cofii2p_b200.model import network as net
    from oracle import restate
    frame = get_frame(0, 4096)
    out = _run(cuda_model, frame, "test")
    idx, xy = net.fine_match(out[4], out[5], out[6])
    ridx, rxy = restate.fine_match(out[4].cpu(), out[5].cpu(), out[6].cpu())
    assert torch.equal(idx.cpu(), ridx) and torch.equal(xy.cpu(), rxy)


def test_engine_graph_matches_forward(cuda_model):
    """The CUDA-graph inference engine (static buffers, B=2) reproduces the eager single-frame forward, also after
    new inputs are uploaded from pinned host memory."""
    from cofii2p_b200.engine import InferenceEngine
    from cofii2p_b200.frames import stack_frames
    from cofii2p_b200 import ops
    ops.set_engine("fp32")
    fa = [get_frame(s, 4096) for s in (0, 1)]
    fb = [get_frame(s, 4096) for s in (1, 0)]
    eng = InferenceEngine(cuda_model, stack_frames(fa), use_graph=True)
    assert eng.graph is not None and eng.launches_per_step > 100
    for frames in (fa, fb):
        nb = eng.upload(eng.host_buffers(stack_frames(frames)))
        assert nb > 0
        eng.run()
        eng.download()
        outs = eng.results()
        for f, o in zip(frames, outs):
            single = _run(cuda_model, f, "val")
            for a, b in zip(o, single):
                if b is None:
                    assert a is None
                else:
                    assert rel_err(a, b) < 1e-6


def test_pipelined_engine_matches_forward(cuda_model):
    """Double-buffered H2D / compute / D2H pipeline returns the same results as the eager forward for every batch."""
    from cofii2p_b200.engine import PipelinedEngine, InferenceEngine
    from cofii2p_b200.frames import stack_frames
    from cofii2p_b200 import ops
    ops.set_engine("fp32")
    sets = [[get_frame(s, 4096) for s in pair] for pair in ((0, 1), (1, 0), (1, 1))]
    pipe = PipelinedEngine(cuda_model, stack_frames(sets[0]), depth=2)
    hosts = [pipe.engines[0].host_buffers(stack_frames(fs)) for fs in sets]
    for i, fs in enumerate(sets):
        nin, nout = pipe.step(hosts[i])
        assert nin > 0 and nout > 0
        outs = pipe.last_results()
        for f, o in zip(fs, outs):
            single = _run(cuda_model, f, "val")
            for a, b in zip(o, single):
                assert (a is None and b is None) or rel_err(a, b) < 1e-6


def test_train_mode_batchnorm_matches_oracle(seeded_sd):
    """model.train(): the decoder's BatchNorm uses batch statistics and updates running stats (reference imagenet.py:381-394)."""
    from cofii2p_b200.model.network import CoFiI2P
    from cofii2p_b200.options import Options_KITTI
    from cofii2p_b200.frames import frame_to
    from cofii2p_b200 import ops
    from oracle import restate
    ops.set_engine("fp32")
    m = CoFiI2P(Options_KITTI())
    m.load_state_dict(seeded_sd, strict=True)
    m = m.cuda().train()
    frame = get_frame(0, 4096)
    sd = {k: v.clone() for k, v in seeded_sd.items()}
    with torch.no_grad():
        ref = restate.forward(sd, frame["pc_data_dict"], frame["img"], frame["fine_center_kpt_coors"], frame["fine_xy"],
                              frame["fine_pc_inline_index"], "train", bn_training=True)
        f = frame_to(frame, "cuda")
        got = m(f["pc_data_dict"], f["img"], f["fine_center_kpt_coors"], f["fine_xy"], f["fine_pc_inline_index"], "train")
    for a, b in zip(got, ref):
        assert (a is None and b is None) or rel_err(a, b) < TOL
    new = m.state_dict()
    for k in ("img_upsample_1.conv.0.bn1.running_mean", "img_upsample_2.conv.1.bn2.running_var",
              "img_upsample_2.conv.0.conv_skip.1.running_mean"):
        assert rel_err(new[k], sd[k]) < 1e-4, k          # F.batch_norm(training=True) updated sd in place
        assert not torch.equal(new[k].cpu(), seeded_sd[k])
    assert int(new["img_upsample_1.conv.0.bn1.num_batches_tracked"]) == 1


def test_other_image_size_nuscenes(seeded_sd):
    """Options_Nuscenes geometry (160 x 320 image, reference data/options.py:73-74): generic (non-TMA-tiled) conv shapes,
    20 x 40 coarse grid, reference-literal border mask."""
    from cofii2p_b200.model.network import CoFiI2P
    from cofii2p_b200.options import Options_KITTI
    from cofii2p_b200.frames import frame_to
    from cofii2p_b200 import ops
    from oracle import restate
    opt = Options_KITTI()
    opt.img_W = 320
    m = CoFiI2P(opt)
    m.load_state_dict(seeded_sd, strict=True)
    m = m.cuda().eval()
    frame = get_frame(2, 4096)
    frame = dict(frame)
    frame["img"] = frame["img"][:, :, :, :320].contiguous()
    k = frame["fine_center_kpt_coors"].clone()
    k[0] = k[0].clamp(max=150)
    frame["fine_center_kpt_coors"] = k
    for engine, tol in (("fp32", TOL), ("tf32", 5e-2)):
        ops.set_engine(engine)
        try:
            for mode in ("val", "test"):
                with torch.no_grad():
                    ref = restate.forward(seeded_sd, frame["pc_data_dict"], frame["img"], frame["fine_center_kpt_coors"],
                                          frame["fine_xy"], frame["fine_pc_inline_index"], mode, img_hw=(160, 320))
                    f = frame_to(frame, "cuda")
                    got = m(f["pc_data_dict"], f["img"], f["fine_center_kpt_coors"], f["fine_xy"],
                            f["fine_pc_inline_index"], mode)
                for a, b in zip(got[:4], ref[:4]):
                    assert rel_err(a, b) < tol, (engine, mode, rel_err(a, b))
                if engine == "fp32":
                    for a, b in zip(got[4:], ref[4:]):
                        assert (a is None and b is None) or (tuple(a.shape) == tuple(b.shape) and rel_err(a, b) < tol)
        finally:
            ops.set_engine("fp32")


def test_graph_cached_forward_matches_eager(cuda_model):
    """enable_cuda_graph(): the per-frame forward replays a cached graph of the static-shape core; results identical."""
    from cofii2p_b200 import ops
    ops.set_engine("fp32")
    frames = [get_frame(s, 4096) for s in (0, 1)]
    eager = [[_run(cuda_model, f, mode) for mode in ("val", "test")] for f in frames]
    cuda_model.enable_cuda_graph(True)
    try:
        for rep in range(2):
            for f, e in zip(frames, eager):
                for mode, ref in zip(("val", "test"), e):
                    got = _run(cuda_model, f, mode)
                    for a, b in zip(got, ref):
                        assert (a is None and b is None) or torch.equal(a, b)
        assert len(cuda_model._graphs) == 1
    finally:
        cuda_model.enable_cuda_graph(False)


@pytest.mark.parametrize("seed,n", [(0, 4096), (1, 4096), (0, 20480)])
def test_pose_matches_reference_pipeline(cuda_model, seed, n):
    """Row f2: forward(test) -> fine matching -> cv2.solvePnPRansac, against the same steps applied to the REAL
    reference's frozen outputs (tests/golden): correspondences identical, RTE/RRE within 1e-3 (north star)."""
    import numpy as np
    from cofii2p_b200 import evaluate as ev, ops
    from cofii2p_b200.frames import frame_to
    from oracle import evaluate as oev
    ops.set_engine("fp32")
    f = get_frame(seed, n)
    z = np.load(os.path.join(ROOT, "tests", "golden", f"frame_s{seed}_n{n}.npz"))
    K = f["K_half"].numpy()
    T_gt = np.linalg.inv(f["P_cloud_from_cam"].numpy().astype(np.float64))
    mine = ev.register(cuda_model, frame_to(f, "cuda"), K, T_gt)
    ri, ro, ridx = oev.correspondences(*[torch.from_numpy(z["test/" + k]) for k in
                                         ("fine_img_feature_patch", "fine_pc_inline_feature", "fine_center_xy", "coarse_pc_points")])
    assert np.array_equal(mine["image_points"], ri) and np.array_equal(mine["object_points"], ro)
    assert torch.equal(mine["fine_index"].cpu(), ridx)
    ok, T_ref, _ = ev.solve_pose(K, ri, ro)
    assert ok == mine["success"]
    if ok:
        rte_r, rre_r = ev.pose_error(T_ref, T_gt)
        assert abs(mine["rte"] - rte_r) <= 1e-3 and abs(mine["rre"] - rre_r) <= 1e-3
        assert np.allclose(mine["T"], T_ref, atol=1e-9)


def test_kitti_front_end_feeds_the_model(cuda_model, tmp_path):
    """Row f4 end to end: a frame read from the reference's on-disk layout (synthetic sequence), pyramid + KNN-128 tables
    built on the GPU by the library, through forward(test) and forward(val) -- the call pattern of evaluation/eval_all.py:60-96."""
    import numpy as np
    from cofii2p_b200 import ops
    from cofii2p_b200.data.kitti import KittiFrames, write_synthetic_sequence
    from cofii2p_b200.frames import frame_to
    from oracle import knn as ok
    sys_path_opt = __import__("test_cpu")._kitti_opt
    root = str(tmp_path)
    write_synthetic_sequence(root, 9, 1, n_points=60000)
    ds = KittiFrames(sys_path_opt(root, 20480), "val")
    item = ds[0]
    d = item["pc_data_dict"]
    assert d["points"][0].is_cuda and d["neighbors"][0].shape == (20480, 128)
    # the device-built tables equal the CPU oracle's on the same pyramid (spot check of one level)
    ref = ok.knn_table(d["points"][3].cpu().numpy(), d["points"][3].cpu().numpy(), 128, ok.DIRECT)
    dist = lambda t: ((d["points"][3].cpu()[:, None, :] - d["points"][3].cpu()[t]) ** 2).sum(-1)
    assert torch.equal(dist(torch.from_numpy(ref)), dist(d["neighbors"][3].cpu()))       # equal up to ties
    ops.set_engine("fp32")
    f = frame_to(item, "cuda")
    n = item["pc_kpt_idx"].numel()
    with torch.no_grad():
        out = cuda_model(f["pc_data_dict"], f["img"].unsqueeze(0), f["fine_center_kpt_coors"], f["fine_xy_coors"].float(),
                         f["fine_pc_inline_index"], "val")
        out_t = cuda_model(f["pc_data_dict"], f["img"].unsqueeze(0), f["fine_center_kpt_coors"], None, None, "test")
    assert out[0].shape == (1, 128, 20, 64) and out[4].shape == (n, 64, 4, 4) and out[5].shape == (n, 64)
    assert out_t[6].shape[0] == 2 and out_t[7].shape[1] == 3 and out_t[6].shape[1] == out_t[7].shape[0] >= 4
    assert all(bool(torch.isfinite(t).all()) for t in out[:6])


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["fp32", "tf32x3"])
def test_decoder_split_equals_concatenated_linear(cuda_model, engine):
    """KPConvFPN._decode evaluates Linear(cat[up(x_c), x_f]) as up(W_c x_c) + W_f x_f (the coarse half at the coarse resolution,
    reference model/kpconv/kp_backbone.py:100-118).  Against the concatenated form run through the same block: decoder4
    (UnaryBlock: Linear + GroupNorm + LeakyReLU, statistics from the accumulate GEMM's epilogue) and decoder2 (Linear only),
    two stacked frames."""
    from cofii2p_b200 import ops
    fpn = cuda_model.pc_encoder
    g = torch.Generator().manual_seed(11)
    frames, nc, nf = 2, 256, 512
    up = torch.randint(0, nc, (frames * nf, 128), generator=g).cuda()
    ops.set_engine(engine)
    try:
        with torch.no_grad():
            for block, c1, c2 in ((fpn.decoder4, 2048, 1024), (fpn.decoder2, 512, 256)):
                xc = torch.randn((frames * nc, c1), generator=g).cuda()
                xf = torch.randn((frames * nf, c2), generator=g).cuda()
                want = block(fpn._up_cat(xc, up, xf, frames), frames)
                got = fpn._decode(block, xc, up, xf, frames)
                assert got.shape == want.shape
                assert rel_err(got, want.cpu()) < 2e-5, (engine, c1, rel_err(got, want.cpu()))
    finally:
        ops.set_engine("fp32")
