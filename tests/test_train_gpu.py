"""Training path: every forward op's hand-written backward (cofii2p_b200/autograd.py -> csrc/backward.cu) against torch
autograd on the CPU oracle formulation, then the whole model's parameter gradients through the reference's three losses
against the oracle's autograd (oracle/restate.py with requires_grad state dict).  Tolerance 2e-4 per op (fp32 engine)."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from conftest import get_frame, rel_err

pytestmark = pytest.mark.gpu
ENGINE = os.environ.get("COFI_TEST_ENGINE", "fp32")      # tf32: same tests at tensor-core tolerance
TOL = 2e-4 if ENGINE == "fp32" else 6e-3


def _ad():
    from cofii2p_b200 import autograd as ad, ops
    ops.set_engine(ENGINE)
    return ad, ops


def _leaf(t):
    return t.clone().requires_grad_(True)


def _cuda_leaf(t):
    return t.cuda().requires_grad_(True)


def _nrm_err(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-30))


def _check(pairs):
    # fp32 engine: element-wise (max-norm) bound.  tf32 engine: a rounded pre-activation within 1e-3 of zero flips its
    # ReLU / max mask, which moves single gradient elements by O(1) -- the bound is on the whole tensor (2-norm) there.
    for name, got, ref in pairs:
        e = rel_err(got, ref) if ENGINE == "fp32" else _nrm_err(got, ref)
        assert e < TOL, (name, e)


@pytest.mark.parametrize("R,Mo,No", [(5000, 960, 64), (40960, 64, 256), (4100, 128, 192), (300, 48, 64), (33, 8, 4)])
def test_gemm_tn_engines(R, Mo, No):
    """dW = A^T B on the tcgen05 MN-major engine and on the SIMT engine against fp64."""
    from cofii2p_b200 import ops
    g = torch.Generator().manual_seed(R)
    a, b = torch.randn((R, Mo), generator=g), torch.randn((R, No), generator=g)
    ref = a.double().t() @ b.double()
    for eng, tol in (("fp32", 2e-5), ("tf32", 2e-3)):
        ops.set_engine(eng)
        got = ops.gemm_tn(a.cuda(), b.cuda())
        assert _nrm_err(got, ref) < tol, (eng, _nrm_err(got, ref))
        assert rel_err(got, ref) < tol * 5, (eng, rel_err(got, ref))
    ops.set_engine(ENGINE)


@pytest.mark.parametrize("act", [0, 1, 2, 3])
def test_linear_backward(act):
    ad, ops = _ad()
    g = torch.Generator().manual_seed(act)
    x, w, b = torch.randn((300, 64), generator=g), torch.randn((48, 64), generator=g) / 8, torch.randn((48,), generator=g)
    go = torch.randn((300, 48), generator=g)
    xr, wr, br = _leaf(x), _leaf(w), _leaf(b)
    y = F.linear(xr, wr, br)
    y = [y, F.relu(y), F.leaky_relu(y, 0.1), torch.sigmoid(y)][act]
    y.backward(go)
    xc, wc, bc = _cuda_leaf(x), _cuda_leaf(w), _cuda_leaf(b)
    yc = ad.linear(xc, wc, bc, act)
    yc.backward(go.cuda())
    if ENGINE != "fp32" and act in (1, 2):
        # tensor-core rounding flips the activation mask of pre-activations within 1e-3 of zero: take the mask the
        # CUDA forward actually produced and check the contractions behind it
        slope = torch.where(yc.detach().cpu() > 0, 1.0, 0.0 if act == 1 else 0.1)
        gm = (go * slope).double()
        _check([("y", yc, y), ("dx", xc.grad, gm @ w.double()), ("dw", wc.grad, gm.t() @ x.double()), ("db", bc.grad, gm.sum(0))])
    else:
        _check([("y", yc, y), ("dx", xc.grad, xr.grad), ("dw", wc.grad, wr.grad), ("db", bc.grad, br.grad)])
    # 1-channel output (the score heads' last layer)
    w1, go1 = w[:1].clone(), go[:, :1].clone()
    xr, wr = _leaf(x), _leaf(w1)
    torch.sigmoid(F.linear(xr, wr)).backward(go1)
    xc, wc = _cuda_leaf(x), _cuda_leaf(w1)
    ad.linear(xc, wc, None, 3).backward(go1.cuda())
    _check([("dx1", xc.grad, xr.grad), ("dw1", wc.grad, wr.grad)])


def test_kpconv_backward():
    ad, ops = _ad()
    from oracle import restate
    g = torch.Generator().manual_seed(1)
    n, c, co, sigma = 600, 32, 32, 0.4
    pts = torch.rand((n, 3), generator=g) * 3
    nbr = torch.cdist(pts.double(), pts.double()).topk(128, dim=1, largest=False).indices
    nbr[:50, 100:] = n  # shadow tail
    feats, w, b = torch.randn((n, c), generator=g), torch.randn((15, c, co), generator=g) / 6, torch.randn((co,), generator=g)
    kp = torch.randn((15, 3), generator=g) * 0.3
    go = torch.randn((n, co), generator=g)
    fr, wr, br = _leaf(feats), _leaf(w), _leaf(b)
    y = restate.kpconv(fr, pts, pts, nbr, wr, br, kp, sigma)
    y.backward(go)
    fc, wc, bc = _cuda_leaf(feats), _cuda_leaf(w), _cuda_leaf(b)
    yc = ad.kpconv(fc, wc, bc, pts.cuda(), pts.cuda(), nbr.cuda(), kp.cuda(), sigma, 1, float(kp.norm(dim=1).max()))
    yc.backward(go.cuda())
    _check([("y", yc, y), ("dfeats", fc.grad, fr.grad), ("dw", wc.grad, wr.grad), ("db", bc.grad, br.grad)])


@pytest.mark.parametrize("rows,c,groups,frames,affine,act,res", [(500, 64, 32, 2, True, 2, True), (700, 128, 128, 1, False, 1, False),
                                                                (640, 32, 32, 1, True, 0, False)])
def test_norm_rows_backward(rows, c, groups, frames, affine, act, res):
    ad, ops = _ad()
    g = torch.Generator().manual_seed(rows + c)
    x = torch.randn((frames * rows, c), generator=g) * 2 + 1
    gamma, beta = torch.randn((c,), generator=g), torch.randn((c,), generator=g)
    r = torch.randn((frames * rows, c), generator=g)
    go = torch.randn((frames * rows, c), generator=g)
    xr, gr, br, rr = _leaf(x), _leaf(gamma), _leaf(beta), _leaf(r)
    outs = []
    for f in range(frames):
        yy = F.group_norm(xr[f * rows:(f + 1) * rows].t().unsqueeze(0), groups, gr if affine else None, br if affine else None,
                          1e-5).squeeze(0).t()
        if res:
            yy = yy + rr[f * rows:(f + 1) * rows]
        outs.append([yy, F.relu(yy), F.leaky_relu(yy, 0.1)][act])
    y = torch.cat(outs)
    y.backward(go)
    xc, gc, bc, rc = _cuda_leaf(x), _cuda_leaf(gamma), _cuda_leaf(beta), _cuda_leaf(r)
    yc = ad.norm_rows(xc, frames, groups, gc if affine else None, bc if affine else None, 1e-5, rc if res else None, act)
    yc.backward(go.cuda())
    pairs = [("y", yc, y), ("dx", xc.grad, xr.grad)]
    if affine:
        pairs += [("dgamma", gc.grad, gr.grad), ("dbeta", bc.grad, br.grad)]
    if res:
        pairs += [("dres", rc.grad, rr.grad)]
    _check(pairs)


def test_row_norms_backward():
    ad, ops = _ad()
    g = torch.Generator().manual_seed(3)
    x = torch.randn((333, 256), generator=g) + 0.3
    gamma, beta, r = torch.randn((256,), generator=g), torch.randn((256,), generator=g), torch.randn((333, 256), generator=g)
    go = torch.randn((333, 256), generator=g)
    for act in (0, 1):
        xr, gr, br, rr = _leaf(x), _leaf(gamma), _leaf(beta), _leaf(r)
        y = F.layer_norm(xr, (256,), gr, br, 1e-5)
        y = (F.relu(y) if act else y) + rr
        y.backward(go)
        xc, gc, bc, rc = _cuda_leaf(x), _cuda_leaf(gamma), _cuda_leaf(beta), _cuda_leaf(r)
        yc = ad.layer_norm(xc, gc, bc, 1e-5, act, rc)
        yc.backward(go.cuda())
        _check([("y", yc, y), ("dx", xc.grad, xr.grad), ("dgamma", gc.grad, gr.grad), ("dbeta", bc.grad, br.grad),
                ("dres", rc.grad, rr.grad)])
    xr, ar = _leaf(x), _leaf(r)
    y = F.normalize(xr, dim=1) + ar
    y.backward(go)
    xc, ac = _cuda_leaf(x), _cuda_leaf(r)
    yc = ad.l2norm(xc, ac)
    yc.backward(go.cuda())
    _check([("l2 y", yc, y), ("l2 dx", xc.grad, xr.grad), ("l2 dadd", ac.grad, ar.grad)])
    q = torch.randn((2 * 320, 128), generator=g)
    gq = torch.randn((2 * 320, 128), generator=g)
    qr = _leaf(q)
    y = torch.cat([F.normalize(qr[:320].view(1, 320, 4, 32)).view(320, 128), F.normalize(qr[320:].view(1, 320, 4, 32)).view(320, 128)])
    y.backward(gq)
    qc = _cuda_leaf(q)
    yc = ad.colnorm(qc, 2)
    yc.backward(gq.cuda())
    _check([("colnorm y", yc, y), ("colnorm dx", qc.grad, qr.grad)])


def test_gather_and_maxpool_backward():
    ad, ops = _ad()
    from oracle import restate
    g = torch.Generator().manual_seed(4)
    x = torch.randn((400, 96), generator=g)
    up = torch.randint(0, 401, (900, 128), generator=g)  # duplicates + shadow index 400
    go = torch.randn((900, 96), generator=g)
    xr = _leaf(x)
    restate.nearest_upsample(xr, up).backward(go)
    xc = _cuda_leaf(x)
    ad.gather(xc, up.cuda(), 128, 1, 900).backward(go.cuda())
    _check([("gather dx", xc.grad, xr.grad)])
    nbr = torch.randint(0, 401, (200, 128), generator=g)
    go2 = torch.randn((200, 96), generator=g)
    xr = _leaf(x)
    restate.maxpool(xr, nbr).backward(go2)
    xc = _cuda_leaf(x)
    ad.maxpool_rows(xc, nbr.cuda(), 1).backward(go2.cuda())
    _check([("maxpool dx", xc.grad, xr.grad)])


def test_attention_backward():
    ad, ops = _ad()
    g = torch.Generator().manual_seed(5)
    L, S, frames = 300, 260, 2
    q, k, v = (torch.randn((frames * n, 128), generator=g) for n in (L, S, S))
    go = torch.randn((frames * L, 128), generator=g)
    qr, kr, vr = _leaf(q), _leaf(k), _leaf(v)
    outs = []
    for f in range(frames):
        qq, kk, vv = qr[f * L:(f + 1) * L].view(1, L, 4, 32), kr[f * S:(f + 1) * S].view(1, S, 4, 32), vr[f * S:(f + 1) * S].view(1, S, 4, 32)
        a = torch.softmax(torch.einsum("nlhd,nshd->nlsh", qq, kk) / 32 ** 0.5, dim=2)
        outs.append(torch.einsum("nlsh,nshd->nlhd", a, vv).reshape(L, 128))
    y = torch.cat(outs)
    y.backward(go)
    qc, kc, vc = _cuda_leaf(q), _cuda_leaf(k), _cuda_leaf(v)
    yc = ad.attention(qc, kc, vc, frames, 4, 1.0 / 32 ** 0.5)
    yc.backward(go.cuda())
    _check([("y", yc, y), ("dq", qc.grad, qr.grad), ("dk", kc.grad, kr.grad), ("dv", vc.grad, vr.grad)])


@pytest.mark.parametrize("cin,cout,k,stride,pad,h,w", [(64, 64, 3, 1, 1, 12, 16), (64, 128, 3, 2, 1, 12, 16), (64, 128, 1, 2, 0, 12, 16),
                                                       (3, 64, 7, 2, 3, 16, 24), (192, 64, 3, 1, 1, 8, 12),
                                                       # output widths that are multiples of 32: the tcgen05 weight-gradient path
                                                       (64, 64, 3, 1, 1, 6, 64), (64, 128, 3, 2, 1, 10, 128), (64, 128, 1, 2, 0, 8, 64),
                                                       (192, 64, 3, 1, 1, 4, 32), (128, 128, 3, 1, 1, 5, 96)])
def test_conv_backward(cin, cout, k, stride, pad, h, w):
    ad, ops = _ad()
    g = torch.Generator().manual_seed(cin + cout + k)
    x = torch.randn((2, cin, h, w), generator=g)
    wt = torch.randn((cout, cin, k, k), generator=g) / math.sqrt(cin * k * k)
    xr, wr = _leaf(x), _leaf(wt)
    y = F.conv2d(xr, wr, None, stride, pad)
    go = torch.randn(y.shape, generator=g)
    y.backward(go)
    cpad = (cin + 3) // 4 * 4
    xn = ops.nchw_to_nhwc(x.cuda(), cpad=cpad).requires_grad_(cin % 4 == 0)
    wc = _cuda_leaf(wt)
    yc = ad.conv2d(xn, wc, stride, pad)
    yc.backward(ops.nchw_to_nhwc(go.cuda()))
    pairs = [("y", ops.nhwc_to_nchw(yc.detach()), y), ("dw", wc.grad, wr.grad)]
    if cin % 4 == 0:
        pairs.append(("dx", ops.nhwc_to_nchw(xn.grad), xr.grad))
    _check(pairs)


def test_image_helpers_backward():
    ad, ops = _ad()
    g = torch.Generator().manual_seed(6)
    x = torch.randn((2, 32, 10, 14), generator=g)
    x2 = torch.randn((2, 16, 20, 28), generator=g)
    xr, x2r = _leaf(x), _leaf(x2)
    y = torch.cat((F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=False), x2r), 1)
    go = torch.randn(y.shape, generator=g)
    y.backward(go)
    xc = ops.nchw_to_nhwc(x.cuda()).requires_grad_(True)
    x2c = ops.nchw_to_nhwc(x2.cuda()).requires_grad_(True)
    ad.upsample2x_cat(xc, x2c).backward(ops.nchw_to_nhwc(go.cuda()))
    _check([("up dx1", ops.nhwc_to_nchw(xc.grad), xr.grad), ("up dx2", ops.nhwc_to_nchw(x2c.grad), x2r.grad)])
    xm = torch.randn((2, 16, 11, 13), generator=g)
    xmr = _leaf(xm)
    ym = F.max_pool2d(xmr, 3, 2, 1)
    gm = torch.randn(ym.shape, generator=g)
    ym.backward(gm)
    xmc = ops.nchw_to_nhwc(xm.cuda()).requires_grad_(True)
    ad.maxpool2d(xmc).backward(ops.nchw_to_nhwc(gm.cuda()))
    _check([("maxpool2d dx", ops.nhwc_to_nchw(xmc.grad), xmr.grad)])
    fmap = torch.randn((1, 64, 80, 256), generator=g)
    ctr = torch.stack([torch.randint(2, 254, (64,), generator=g), torch.randint(2, 78, (64,), generator=g)]).float()
    fr = _leaf(fmap)
    from oracle import restate
    p = torch.squeeze(restate.extract_patch(fr, ctr))
    gp = torch.randn(p.shape, generator=g)
    p.backward(gp)
    fc = ops.nchw_to_nhwc(fmap.cuda()).requires_grad_(True)
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    ad.extract_patch(fc, 0, ctr.cuda(), err).backward(gp.cuda())
    _check([("patch dmap", ops.nhwc_to_nchw(fc.grad), fr.grad)])


def test_adam_kernel_matches_torch():
    ad, ops = _ad()
    g = torch.Generator().manual_seed(7)
    p0 = torch.randn((1000,), generator=g)
    pr = _leaf(p0)
    opt = torch.optim.Adam([pr], lr=1e-3)
    pc, m, v = p0.cuda().clone(), torch.zeros(1000, device="cuda"), torch.zeros(1000, device="cuda")
    for step in range(1, 4):
        grad = torch.randn((1000,), generator=g)
        pr.grad = grad.clone()
        opt.step()
        ops.adam_step(pc, grad.cuda(), m, v, 1e-3, 0.9, 0.999, 1e-8, step)
    assert rel_err(pc, pr.detach()) < 1e-6


# ---------------------------------------------------------------------------------------------- whole model
def _fresh_model(seeded_sd):
    from cofii2p_b200.model.network import CoFiI2P
    from cofii2p_b200.options import Options_KITTI
    opt = Options_KITTI()
    m = CoFiI2P(opt)
    m.load_state_dict(seeded_sd, strict=True)
    return m.cuda(), opt


def _sup(frame):
    return {k: frame[k] for k in ("pc_kpt_idx", "pc_outline_idx", "coarse_img_kpt_idx", "K_4", "P", "fine_xy",
                                  "fine_center_kpt_coors")}


def test_model_gradients_vs_oracle_autograd(seeded_sd):
    """loss and every parameter gradient of one train-mode frame against torch autograd through the CPU oracle."""
    from cofii2p_b200 import ops
    from cofii2p_b200.frames import frame_to, stack_frames
    from cofii2p_b200.train import TrainStep, training_losses
    from oracle import restate
    ops.set_engine("fp32")
    frame = get_frame(0, 4096)
    model, opt = _fresh_model(seeded_sd)
    ts = TrainStep(model, opt)
    loss, parts = ts.backward(stack_frames([frame_to(frame, "cuda")]))

    sdr = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k and "kernel_points" not in k
               else v.clone()) for k, v in seeded_sd.items()}
    out = restate.forward(sdr, frame["pc_data_dict"], frame["img"], frame["fine_center_kpt_coors"], frame["fine_xy"],
                          frame["fine_pc_inline_index"], "train", bn_training=True)
    ref_loss, ref_parts, _ = training_losses(out, _sup(frame), opt, frame["pc_data_dict"]["points"][-1])
    ref_loss.backward()
    ref_loss = ref_loss.detach()
    assert abs(float(loss) - float(ref_loss)) < 1e-4 * max(1.0, abs(float(ref_loss))), (float(loss), float(ref_loss))
    assert rel_err(parts, torch.stack(ref_parts)) < 1e-4
    live_ref = {k for k, v in sdr.items() if v.requires_grad and v.grad is not None and float(v.grad.abs().max()) > 0}
    got = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    assert live_ref <= set(got), sorted(live_ref - set(got))[:8]
    worst = ("", 0.0)
    for k in sorted(got):
        g_ref = sdr[k].grad
        if g_ref is None:
            assert float(got[k].abs().max()) == 0.0, k
            continue
        num, den = float((got[k].cpu().double() - g_ref.double()).norm()), float(g_ref.double().norm())
        if den < 1e-4:
            # a bias that feeds a per-channel normalisation has a mathematically zero gradient: both sides hold 1e-5-size
            # rounding noise (live gradients have norms 1e-3 .. 1e+1)
            assert float(got[k].norm()) < 1e-4, k
            continue
        e = num / den
        if e > worst[1]:
            worst = (k, e)
    print("worst gradient:", worst, "live tensors:", len(live_ref))
    assert worst[1] < 2e-3, worst


def test_stacked_frames_average_the_per_frame_gradients(seeded_sd):
    from cofii2p_b200 import ops
    from cofii2p_b200.frames import frame_to, stack_frames
    from cofii2p_b200.train import TrainStep
    ops.set_engine("fp32")
    frames = [frame_to(get_frame(s, 4096), "cuda") for s in (0, 1)]
    grads = []
    for fs in ([frames[0]], [frames[1]], frames):
        model, opt = _fresh_model(seeded_sd)
        ts = TrainStep(model, opt)
        ts.backward(stack_frames(fs))
        grads.append(ts.flat_g.clone())
    avg = 0.5 * (grads[0] + grads[1])
    assert float((grads[2] - avg).norm() / avg.norm()) < 1e-4


def test_first_adam_step_moves_by_lr(seeded_sd):
    from cofii2p_b200 import ops
    from cofii2p_b200.frames import frame_to, stack_frames
    from cofii2p_b200.train import TrainStep
    ops.set_engine("fp32")
    model, opt = _fresh_model(seeded_sd)
    ts = TrainStep(model, opt)
    batch = stack_frames([frame_to(get_frame(0, 4096), "cuda")])
    l0, _ = ts.step(batch)
    p_after, g = ts.flat_p.clone(), ts.flat_g.clone()
    # step 1 of Adam: p -= lr * g / (|g| + eps)
    want = -opt.lr * g / (g.abs() + 1e-8)
    for name, p in model.named_parameters():
        if p.grad is not None:
            assert rel_err(p.detach().cpu(), (seeded_sd[name] + want[_offset(ts, p):_offset(ts, p) + p.numel()].view_as(p).cpu())) < 1e-5
            break
    losses = [float(l0)] + [float(ts.step(batch)[0]) for _ in range(5)]
    assert all(math.isfinite(x) for x in losses)
    assert losses[-1] < losses[0], losses


def _offset(ts, p):
    return (p.data_ptr() - ts.flat_p.data_ptr()) // 4


def test_captured_step_equals_eager_step(seeded_sd):
    """TrainStep.enable_cuda_graph: the replayed forward + losses + backward gives the eager step's loss and gradient
    (the backward's vector atomics reorder fp32 sums, hence a tolerance), also for a batch other than the captured one."""
    from cofii2p_b200 import ops
    from cofii2p_b200.frames import frame_to, stack_frames
    from cofii2p_b200.train import TrainStep
    ops.set_engine("fp32")
    b0 = stack_frames([frame_to(get_frame(0, 4096), "cuda")])
    b1 = stack_frames([frame_to(get_frame(1, 4096), "cuda")])
    model, opt = _fresh_model(seeded_sd)
    eager = TrainStep(model, opt)
    ref = {}
    for name, b in (("b0", b0), ("b1", b1)):
        l, _ = eager.backward(b)
        ref[name] = (float(l), eager.flat_g.clone())
    model2, _ = _fresh_model(seeded_sd)
    ts = TrainStep(model2, opt, lr=0.0)                       # lr 0: parameters stay put, steps are comparable
    keep = stack_frames([frame_to(get_frame(0, 4096), "cuda")])
    ts.enable_cuda_graph(keep)
    for name, b in (("b0", keep), ("b1", b1), ("b0", b0)):
        l, _ = ts.step(b)
        ts.check_errors()
        assert abs(float(l) - ref[name][0]) < 1e-4 * abs(ref[name][0]), name
        assert float((ts.flat_g - ref[name][1]).norm() / ref[name][1].norm()) < 1e-4, name


@pytest.mark.parametrize("L,S,frames", [(300, 260, 2), (1280, 1280, 1), (64, 132, 3)])
def test_attention_backward_tcgen05(L, S, frames):
    """cofi_attention_bwd_tc (tf32 operands) against torch autograd in fp32, ragged tiles and stacked frames included."""
    from cofii2p_b200 import autograd as ad, ops
    ops.set_engine("tf32")
    try:
        g = torch.Generator().manual_seed(L + S)
        q, k, v = (torch.randn((frames * n, 128), generator=g) for n in (L, S, S))
        go = torch.randn((frames * L, 128), generator=g)
        qr, kr, vr = _leaf(q), _leaf(k), _leaf(v)
        outs = []
        for f in range(frames):
            qq, kk, vv = (qr[f * L:(f + 1) * L].view(1, L, 4, 32), kr[f * S:(f + 1) * S].view(1, S, 4, 32),
                          vr[f * S:(f + 1) * S].view(1, S, 4, 32))
            a = torch.softmax(torch.einsum("nlhd,nshd->nlsh", qq, kk) / 32 ** 0.5, dim=2)
            outs.append(torch.einsum("nlsh,nshd->nlhd", a, vv).reshape(L, 128))
        y = torch.cat(outs)
        y.backward(go)
        qc, kc, vc = _cuda_leaf(q), _cuda_leaf(k), _cuda_leaf(v)
        yc = ad.attention(qc, kc, vc, frames, 4, 1.0 / 32 ** 0.5)
        yc.backward(go.cuda())
        for name, got, ref in (("y", yc, y), ("dq", qc.grad, qr.grad), ("dk", kc.grad, kr.grad), ("dv", vc.grad, vr.grad)):
            assert _nrm_err(got, ref) < 3e-3, (name, _nrm_err(got, ref))
            assert rel_err(got, ref) < 2e-2, (name, rel_err(got, ref))
        # and against the SIMT fp32 backward of the same library on the same inputs
        ops.set_engine("fp32")
        q2, k2, v2 = _cuda_leaf(q), _cuda_leaf(k), _cuda_leaf(v)
        ad.attention(q2, k2, v2, frames, 4, 1.0 / 32 ** 0.5).backward(go.cuda())
        for name, got, ref in (("dq", qc.grad, q2.grad), ("dk", kc.grad, k2.grad), ("dv", vc.grad, v2.grad)):
            assert _nrm_err(got, ref) < 3e-3, (name, _nrm_err(got, ref))
    finally:
        ops.set_engine("fp32")


@pytest.mark.parametrize("captured", [False, True])
def test_eval_after_training_uses_the_updated_weights(seeded_sd, captured):
    """The fused Adam kernel writes parameters through raw pointers (no tensor._version bump) and a replayed step graph
    updates BatchNorm running statistics the same way: the eval-path caches (K-major KPConv / conv weight packs, folded
    BatchNorm, captured inference graphs) must be invalidated by the step (ops.weights_epoch).  Train k steps, then
    compare model.eval() -- eager, graph-cached forward and a pre-built InferenceEngine -- against a FRESH model loaded
    from the trained state_dict (the reference validates inside its training loop, train.py:70)."""
    from cofii2p_b200 import ops
    from cofii2p_b200.engine import InferenceEngine
    from cofii2p_b200.frames import frame_to, stack_frames
    from cofii2p_b200.model.network import CoFiI2P
    from cofii2p_b200.options import Options_KITTI
    from cofii2p_b200.train import TrainStep
    ops.set_engine("fp32")
    frame = frame_to(get_frame(0, 4096), "cuda")
    batch = stack_frames([frame])
    args = [frame[k] for k in ("pc_data_dict", "img", "fine_center_kpt_coors", "fine_xy", "fine_pc_inline_index")]
    model, opt = _fresh_model(seeded_sd)
    model.eval()
    with torch.no_grad():
        before = model(*args, "val")                      # fills every eval-path cache with the initial weights
    model.enable_cuda_graph(True)
    with torch.no_grad():
        model(*args, "val")                               # ... and captures a graph of them
    eng = InferenceEngine(model, batch, mode="val", use_graph=True)
    ts = TrainStep(model, opt, lr=1e-3)
    if captured:
        ts.enable_cuda_graph(stack_frames([frame_to(get_frame(0, 4096), "cuda")]))
    for _ in range(3):
        ts.step(batch)
    fresh = CoFiI2P(Options_KITTI())
    fresh.load_state_dict({k: v.detach().cpu().clone() for k, v in model.state_dict().items()}, strict=True)
    fresh = fresh.cuda().eval()
    model.eval()
    with torch.no_grad():
        want = fresh(*args, "val")
        got_graph = model(*args, "val")
        model.enable_cuda_graph(False)
        got_eager = model(*args, "val")
    eng.run()
    got_engine = eng.results()[0]
    assert rel_err(want[0], before[0]) > 1e-4             # the three steps really moved the weights
    for got in (got_graph, got_eager, got_engine):
        for a, b in zip(got[:6], want[:6]):
            assert rel_err(a, b) < 1e-5, rel_err(a, b)


# ---------------------------------------------------------------------------------------------- fused losses (row f3)
def test_fused_loss_kernels_match_the_reference_formulas():
    """csrc/loss.cu (forward + analytic backward in one launch) against torch autograd through the reference's formulas
    (model/loss.py:9-93 restated in cofii2p_b200/model/loss.py::*_algebra, bit-exact vs the reference on CPU)."""
    from cofii2p_b200.model import loss as L
    g = torch.Generator().manual_seed(5)
    n, C = 64, 128
    img = F.normalize(torch.randn((C, n), generator=g), dim=0)
    pc = F.normalize(img + 0.7 * torch.randn((C, n), generator=g), dim=0)
    mask = (torch.rand((n, n), generator=g) < 0.04).float()
    mask[torch.arange(n), torch.arange(n)] = 1.0
    a_ref, b_ref = _leaf(img), _leaf(pc)
    l_ref, d_ref = L.desc_loss_algebra("cpu", a_ref, b_ref, mask, pos_margin=0.2, neg_margin=1.8)
    l_ref.backward()
    a, b = _cuda_leaf(img), _cuda_leaf(pc)
    l, d = L.desc_loss("cuda", a, b, mask.cuda(), pos_margin=0.2, neg_margin=1.8)
    (l * 1.0).backward()
    assert abs(float(l) - float(l_ref)) < 1e-5 * abs(float(l_ref)) and rel_err(d, d_ref) < 1e-6
    assert _nrm_err(a.grad, a_ref.grad) < 1e-5 and _nrm_err(b.grad, b_ref.grad) < 1e-5
    # overlap (BCE), including a saturated score
    si, so = torch.rand((64,), generator=g) * 0.98 + 0.01, torch.rand((64,), generator=g) * 0.98 + 0.01
    si[0], so[0] = 1.0 - 1e-7, 1e-7
    si_r, so_r = _leaf(si), _leaf(so)
    lo_ref = L.overlap_loss_algebra("cpu", si_r, so_r)
    lo_ref.backward()
    si_c, so_c = _cuda_leaf(si), _cuda_leaf(so)
    lo = L.overlap_loss("cuda", si_c, so_c)
    lo.backward()
    assert abs(float(lo) - float(lo_ref)) < 1e-5 * abs(float(lo_ref))
    assert _nrm_err(si_c.grad, si_r.grad) < 1e-5 and _nrm_err(so_c.grad, so_r.grad) < 1e-5
    # fine circle loss
    patch = F.normalize(torch.randn((64, 64, 4, 4), generator=g), dim=1)
    fpc = F.normalize(torch.randn((64, 64), generator=g) + 2.0 * patch[:, :, 1, 2], dim=1)
    rel = torch.randint(0, 16, (64,), generator=g)
    p_r, f_r = _leaf(patch), _leaf(fpc)
    lf_ref = L.fine_circle_loss_algebra("cpu", p_r, f_r, rel, 64)
    lf_ref.backward()
    p_c, f_c = _cuda_leaf(patch), _cuda_leaf(fpc)
    lf = L.fine_circle_loss("cuda", p_c, f_c, rel.cuda(), 64)
    lf.backward()
    assert abs(float(lf) - float(lf_ref)) < 1e-5 * abs(float(lf_ref))
    assert _nrm_err(p_c.grad, p_r.grad) < 1e-5 and _nrm_err(f_c.grad, f_r.grad) < 1e-5


def test_fused_training_losses_equal_the_per_frame_path(seeded_sd):
    """TrainStep(fused_losses=True) -- token-layout outputs, batched fused loss kernels -- gives the loss and the gradient of
    TrainStep(fused_losses=False) -- the reference's per-frame formulas on the public outputs through torch autograd."""
    from cofii2p_b200 import ops
    from cofii2p_b200.frames import frame_to, stack_frames
    from cofii2p_b200.train import TrainStep
    ops.set_engine("fp32")
    batch = stack_frames([frame_to(get_frame(s, 4096), "cuda") for s in (0, 1)])
    res = {}
    for fused in (False, True):
        model, opt = _fresh_model(seeded_sd)
        ts = TrainStep(model, opt, fused_losses=fused)
        l, parts = ts.backward(batch)
        res[fused] = (float(l), parts.clone(), ts.flat_g.clone())
    assert abs(res[True][0] - res[False][0]) < 1e-5 * abs(res[False][0])
    assert rel_err(res[True][1], res[False][1]) < 1e-5
    assert float((res[True][2] - res[False][2]).norm() / res[False][2].norm()) < 1e-4
