"""Per-op parity of the CUDA kernels (through the C ABI) against the CPU oracle (oracle/restate.py),
same seeded inputs, sizes the oracle finishes in seconds.  fp32 engine: rel <= 1e-4 per op (the north star's
end-to-end tolerance is 1e-3); integer outputs bit-exact."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _ops():
    from cofii2p_b200 import ops
    ops.set_engine("fp32")
    return ops


def _cloud(g, n, extent):
    return torch.rand((n, 3), generator=g) * extent


def _knn(q, s, k, shadow_frac=0.0, g=None):
    d = torch.cdist(q.double(), s.double())
    idx = d.topk(min(k, s.shape[0]), dim=1, largest=False).indices
    if idx.shape[1] < k:
        idx = torch.cat([idx, torch.full((q.shape[0], k - idx.shape[1]), s.shape[0], dtype=torch.int64)], 1)
    if shadow_frac > 0:  # ragged neighbourhoods: tail entries replaced by the shadow index
        cut = torch.randint(1, k, (q.shape[0], 1), generator=g)
        mask = torch.arange(k)[None, :] >= cut
        drop = torch.rand((q.shape[0], 1), generator=g) < shadow_frac
        idx = torch.where(mask & drop, torch.full_like(idx, s.shape[0]), idx)
    return idx


@pytest.mark.parametrize("cin,cout,n,m,h,sigma,extent,shadow", [
    (4, 64, 1500, 1500, 128, 0.2, 3.0, 0.0),
    (32, 32, 1500, 1500, 128, 0.2, 3.0, 0.3),
    (64, 64, 1024, 512, 128, 0.4, 4.0, 0.0),
    (128, 128, 700, 350, 128, 0.8, 6.0, 0.2),
    (256, 256, 300, 300, 128, 1.6, 8.0, 0.0),
    (512, 512, 200, 200, 128, 3.2, 12.0, 0.0),
    (32, 32, 100, 100, 128, 0.2, 1.0, 0.0),   # fewer points than neighbours: shadow tail
    (32, 32, 600, 600, 37, 0.3, 2.0, 0.0),    # H not a multiple of 32
])
def test_kpconv(cin, cout, n, m, h, sigma, extent, shadow):
    ops = _ops()
    from oracle import restate
    g = torch.Generator().manual_seed(cin * 7 + n)
    s_pts = _cloud(g, n, extent)
    q_pts = s_pts[:m].clone() if m <= n else _cloud(g, m, extent)
    nbr = _knn(q_pts, s_pts, h, shadow, g)
    feats = torch.randn((n, cin), generator=g)
    w = torch.randn((15, cin, cout), generator=g) / math.sqrt(cin)
    b = torch.randn((cout,), generator=g) * 0.1
    kp = torch.randn((15, 3), generator=g) * sigma * 0.8
    kp[0] = 0
    ref = restate.kpconv(feats, q_pts, s_pts, nbr, w, b, kp, sigma)
    packed = ops.pack_points(s_pts.cuda(), feats.cuda())
    agg, cnt = ops.kpconv_aggregate(feats.cuda(), packed, q_pts.cuda(), nbr.cuda(), kp.cuda(), sigma, 1)
    # the exact far-neighbour cull must not change a single bit
    agg2, cnt2 = ops.kpconv_aggregate(feats.cuda(), packed, q_pts.cuda(), nbr.cuda(), kp.cuda(), sigma, 1,
                                      kp_reach=float(kp.norm(dim=1).max()))
    assert torch.equal(agg, agg2) and torch.equal(cnt, cnt2)
    wt = w.reshape(-1, cout).t().contiguous().cuda()
    out = ops.gemm(agg, wt, bias=b.cuda(), rowdiv=cnt)
    assert rel_err(out, ref) < TOL


def test_kpconv_two_frames_equals_two_calls():
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    outs, packs = [], []
    n, c, h = 800, 64, 128
    data = []
    for f in range(2):
        s = _cloud(g, n, 3.0)
        nbr = _knn(s, s, h)
        feats = torch.randn((n, c), generator=g)
        data.append((s, nbr, feats))
    kp = torch.randn((15, 3), generator=g) * 0.3
    for s, nbr, feats in data:
        pk = ops.pack_points(s.cuda(), feats.cuda())
        outs.append(ops.kpconv_aggregate(feats.cuda(), pk, s.cuda(), nbr.cuda(), kp.cuda(), 0.4, 1))
    S = torch.cat([d[0] for d in data]).cuda()
    NB = torch.cat([d[1] for d in data]).cuda()
    FE = torch.cat([d[2] for d in data]).cuda()
    pk = ops.pack_points(S, FE)
    agg, cnt = ops.kpconv_aggregate(FE, pk, S, NB, kp.cuda(), 0.4, 2)
    assert torch.equal(agg, torch.cat([o[0] for o in outs]))
    assert torch.equal(cnt, torch.cat([o[1] for o in outs]))


def test_maxpool_and_upsample():
    ops = _ops()
    from oracle import restate
    g = torch.Generator().manual_seed(1)
    s = _cloud(g, 900, 3.0)
    q = s[:450]
    nbr = _knn(q, s, 128, 0.3, g)
    x = torch.randn((900, 256), generator=g)
    assert torch.equal(ops.maxpool_rows(x.cuda(), nbr.cuda()).cpu(), restate.maxpool(x, nbr))
    up = _knn(s, q, 128)
    xc = torch.randn((450, 96), generator=g)
    got = ops.gather_rows(xc.cuda(), up.cuda(), idx_stride=128, rows_out=900)
    assert torch.equal(got.cpu(), restate.nearest_upsample(xc, up))


@pytest.mark.parametrize("c", [64, 128, 256, 512, 96])
@pytest.mark.parametrize("H", [128, 100, 33])
def test_maxpool_lean_kernels(c, H):
    """The instruction-lean max-pool kernels (fp32: C = 64 / multiples of 128; fp16: C = 64, 128, multiples of 256) and the
    generic fall-back (C = 96) against the oracle: shadow neighbours (a row of zeros must win over negative features), fewer
    than 128 neighbours, negative-only rows, two stacked frames with frame-local indices."""
    ops = _ops()
    from oracle import restate
    g = torch.Generator().manual_seed(c * 7 + H)
    frames, ns, nq = 2, 500, 260
    x = torch.randn((frames * ns, c), generator=g)
    x[: ns // 3] = -x[: ns // 3].abs() - 0.5
    nbr = torch.randint(0, ns, (frames * nq, H), generator=g)
    nbr[::5, -3:] = ns                      # shadow neighbours
    nbr[1::9, 0] = ns                       # ... also in slot 0 (the slot the kernels replicate past H)
    want = torch.cat([restate.maxpool(x[f * ns:(f + 1) * ns], nbr[f * nq:(f + 1) * nq]) for f in range(frames)])
    got = ops.maxpool_rows(x.cuda(), nbr.cuda(), frames)
    assert torch.equal(got.cpu(), want)
    if c % 8 == 0:
        xh = x.cuda().to(torch.float16)
        want_h = torch.cat([restate.maxpool(xh.float().cpu()[f * ns:(f + 1) * ns], nbr[f * nq:(f + 1) * nq]) for f in range(frames)])
        assert torch.equal(ops.maxpool_rows_f16(xh, nbr.cuda(), frames).cpu(), want_h)


@pytest.mark.parametrize("rows,c,groups,frames", [(1000, 64, 32, 1), (777, 128, 32, 2), (640, 2048, 32, 1),
                                                  (1280, 64, 64, 1)])
def test_norm_rows(rows, c, groups, frames):
    ops = _ops()
    g = torch.Generator().manual_seed(rows + c)
    x = torch.randn((frames * rows, c), generator=g) * 2 + 3.0
    gamma = torch.randn((c,), generator=g)
    beta = torch.randn((c,), generator=g)
    res = torch.randn((frames * rows, c), generator=g)
    refs = []
    for f in range(frames):
        xf = x[f * rows:(f + 1) * rows]
        y = F.group_norm(xf.t().unsqueeze(0), groups, gamma, beta, 1e-5).squeeze(0).t()
        refs.append(F.leaky_relu(y + res[f * rows:(f + 1) * rows], 0.1))
    got = ops.norm_rows(x.cuda(), frames, groups, gamma.cuda(), beta.cuda(), 1e-5, residual=res.cuda(), act=ops.ACT_LRELU)
    assert rel_err(got, torch.cat(refs)) < TOL
    # affine-free instance norm (groups == channels)
    got = ops.norm_rows(x.cuda(), frames, c, None, None, 1e-5, act=ops.ACT_RELU)
    refs = [F.relu(F.instance_norm(x[f * rows:(f + 1) * rows].t().unsqueeze(0)).squeeze(0).t()) for f in range(frames)]
    assert rel_err(got, torch.cat(refs)) < TOL


def test_layernorm_l2norm_colnorm():
    ops = _ops()
    g = torch.Generator().manual_seed(2)
    x = torch.randn((333, 1024), generator=g) + 0.5
    gamma, beta = torch.randn((1024,), generator=g), torch.randn((1024,), generator=g)
    res = torch.randn((333, 1024), generator=g)
    ref = F.relu(F.layer_norm(x, (1024,), gamma, beta, 1e-5)) + res
    got = ops.layer_norm_rows(x.cuda(), gamma.cuda(), beta.cuda(), 1e-5, act=ops.ACT_RELU, residual=res.cuda())
    assert rel_err(got, ref) < TOL
    add = torch.randn((333, 1024), generator=g)
    assert rel_err(ops.l2norm_rows(x.cuda(), add=add.cuda()), F.normalize(x, dim=1) + add) < TOL
    q = torch.randn((2 * 640, 128), generator=g)
    ref = torch.cat([F.normalize(q[:640].view(1, 640, 4, 32)).view(640, 128),
                     F.normalize(q[640:].view(1, 640, 4, 32)).view(640, 128)])
    assert rel_err(ops.colnorm_rows(q.cuda(), 2), ref) < TOL


@pytest.mark.parametrize("m,n,k", [(1000, 64, 60), (513, 32, 480), (2000, 128, 32), (300, 1024, 3072), (1280, 1, 64),
                                   (77, 200, 256)])
def test_gemm(m, n, k):
    ops = _ops()
    g = torch.Generator().manual_seed(m + n + k)
    a = torch.randn((m, k), generator=g)
    w = torch.randn((n, k), generator=g) / math.sqrt(k)
    b = torch.randn((n,), generator=g)
    rd = torch.randint(1, 60, (m,), generator=g).float()
    ref = F.leaky_relu(F.linear(a, w) / rd[:, None] + b, 0.1)
    got = ops.gemm(a.cuda(), w.cuda(), bias=b.cuda(), rowdiv=rd.cuda(), act=ops.ACT_LRELU)
    assert rel_err(got, ref) < TOL
    got2 = ops.gemm(a.cuda(), w.cuda(), out=got.clone(), accumulate=True)
    assert rel_err(got2, ref + F.linear(a, w)) < TOL


@pytest.mark.parametrize("cin,cout,k,stride,pad,h,w", [(3, 64, 7, 2, 3, 40, 64), (64, 64, 3, 1, 1, 20, 32),
                                                       (64, 128, 3, 2, 1, 20, 32), (64, 128, 1, 2, 0, 20, 32),
                                                       (192, 128, 3, 1, 1, 10, 24)])
def test_conv2d(cin, cout, k, stride, pad, h, w):
    ops = _ops()
    g = torch.Generator().manual_seed(cin + cout + k)
    x = torch.randn((2, cin, h, w), generator=g)
    wt = torch.randn((cout, cin, k, k), generator=g) / math.sqrt(cin * k * k)
    sc, sh = torch.randn((cout,), generator=g), torch.randn((cout,), generator=g)
    ref_c = F.conv2d(x, wt, None, stride, pad)
    res = torch.randn(ref_c.shape, generator=g)
    ref = F.relu(ref_c * sc[None, :, None, None] + sh[None, :, None, None] + res)
    cpad = (cin + 3) // 4 * 4
    xn = ops.nchw_to_nhwc(x.cuda(), cpad=cpad)
    wp = F.pad(wt.permute(0, 2, 3, 1), (0, cpad - cin)).reshape(cout, -1).contiguous().cuda()
    resn = ops.nchw_to_nhwc(res.cuda())
    y = ops.conv2d_nhwc(xn, wp, k, k, stride, pad, scale=sc.cuda(), shift=sh.cuda(), residual=resn, act=ops.ACT_RELU)
    assert rel_err(ops.nhwc_to_nchw(y), ref) < TOL


def test_image_helpers():
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    x = torch.randn((2, 64, 21, 33), generator=g)
    xn = ops.nchw_to_nhwc(x.cuda())
    assert torch.equal(ops.nhwc_to_nchw(xn).cpu(), x)
    assert torch.equal(ops.nhwc_to_nchw(ops.maxpool2d_3x3s2_nhwc(xn)).cpu(), F.max_pool2d(x, 3, 2, 1))
    x2 = torch.randn((2, 32, 42, 66), generator=g)
    ref = torch.cat((F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False), x2), 1)
    got = ops.nhwc_to_nchw(ops.upsample2x_cat_nhwc(xn, ops.nchw_to_nhwc(x2.cuda())))
    assert rel_err(got, ref) < 1e-6


def test_posenc():
    ops = _ops()
    from oracle import restate
    from cofii2p_b200.model.transformer.position_encoding import PositionEmbeddingCoordsSine
    gy, gx = torch.meshgrid(torch.arange(0, 20), torch.arange(0, 64), indexing="ij")
    xy = torch.stack([gy, gx], -1).reshape(1, -1, 2)
    ref = restate.posenc_sine(xy)[0]
    got = PositionEmbeddingCoordsSine(2, 128)(xy[0].float().cuda())
    assert float((got.cpu() - ref).abs().max()) < 2e-5   # |arg| up to ~400 rad: 1 ulp of the argument
    g = torch.Generator().manual_seed(4)
    pts = (torch.rand((1280, 3), generator=g) - 0.5) * 160
    ref = restate.posenc_sine(pts.unsqueeze(0))[0]
    got = PositionEmbeddingCoordsSine(3, 128)(pts.cuda())
    assert float((got.cpu() - ref).abs().max()) < 1e-4   # |arg| up to ~500 rad
    assert torch.equal(got[:, 126:].cpu(), torch.zeros(1280, 2))


@pytest.mark.parametrize("L,S,frames", [(1280, 1280, 1), (300, 515, 2), (130, 64, 1)])
def test_attention(L, S, frames):
    ops = _ops()
    g = torch.Generator().manual_seed(L + S)
    q = torch.randn((frames * L, 128), generator=g)
    k = torch.randn((frames * S, 128), generator=g)
    v = torch.randn((frames * S, 128), generator=g)
    refs = []
    for f in range(frames):
        qq = q[f * L:(f + 1) * L].view(1, L, 4, 32)
        kk = k[f * S:(f + 1) * S].view(1, S, 4, 32)
        vv = v[f * S:(f + 1) * S].view(1, S, 4, 32)
        a = torch.softmax(torch.einsum("nlhd,nshd->nlsh", qq, kk) / 32 ** 0.5, dim=2)
        refs.append(torch.einsum("nlsh,nshd->nlhd", a, vv).reshape(L, 128))
    got = ops.attention(q.cuda(), k.cuda(), v.cuda(), frames, 4, 1.0 / 32 ** 0.5)
    assert rel_err(got, torch.cat(refs)) < TOL


def test_loftr_layer_and_transformer(seeded_sd, cuda_model):
    _ops()
    from oracle import restate
    g = torch.Generator().manual_seed(6)
    f0 = torch.randn((1, 640, 128), generator=g)
    f1 = torch.randn((1, 400, 128), generator=g)
    ref = restate.loftr_layer(seeded_sd, "transformer.layers.1", f0, f1)
    with torch.no_grad():
        got = cuda_model.transformer.layers[1](f0[0].cuda(), f1[0].cuda(), 1)
    assert rel_err(got, ref[0]) < TOL
    r0, r1 = restate.local_feature_transformer(seeded_sd, f0, f1)
    with torch.no_grad():
        g0, g1 = cuda_model.transformer(f0.cuda(), f1.cuda())
    assert rel_err(g0, r0) < 5e-4 and rel_err(g1, r1) < 5e-4


def test_matching_exact():
    ops = _ops()
    from oracle import restate
    from cofii2p_b200.model import network as net
    g = torch.Generator().manual_seed(7)
    img = F.normalize(torch.randn((1, 128, 20, 64), generator=g), dim=1)
    pc = F.normalize(torch.randn((128, 1280), generator=g), dim=0)
    score = torch.rand((1, 1, 1280), generator=g)
    for thr in (0.9, 0.5, 0.0):
        rxy, ridx = restate.fine_process(score, pc, img, thrs=thr)
        gxy, gidx = net.fine_process(score.cuda(), pc.cuda(), img.cuda(), thrs=thr)
        assert torch.equal(gidx.cpu(), ridx) and torch.equal(gxy.cpu(), rxy)
    # raw arg-min over all points
    px = img.flatten(2)[0].t().contiguous()
    best, val = ops.sim_argmin(pc.t().contiguous().cuda(), px.cuda(), 1)
    dist = 1 - torch.sum(img.flatten(2)[0].unsqueeze(-1) * pc.unsqueeze(-2), dim=0)
    assert torch.equal(best.cpu(), torch.argmin(dist, dim=0))
    assert rel_err(val, dist.min(0).values) < 1e-6
    # point2node
    nodes = (torch.rand((10240, 3), generator=g) - 0.5) * 160
    pts = nodes[torch.randperm(10240, generator=g)[:300]] + 0.01 * torch.randn((300, 3), generator=g)
    ridx = restate.point2node(nodes, pts)
    gidx = net.point2node(nodes.cuda(), pts.cuda()).cpu()
    assert torch.equal(nodes[gidx], nodes[ridx])  # compare gathered coordinates: ties are legitimate
    # extract_patch / fine_match
    fmap = torch.randn((1, 64, 80, 256), generator=g)
    ctr = torch.stack([torch.randint(2, 254, (64,), generator=g), torch.randint(2, 78, (64,), generator=g)]).int()
    rp = torch.squeeze(restate.extract_patch(fmap, ctr))
    gp = torch.squeeze(net.extract_patch(fmap.cuda(), ctr.cuda()))
    assert torch.equal(gp.cpu(), rp)
    pcf = F.normalize(torch.randn((64, 64), generator=g), dim=1)
    ridx, rxy = restate.fine_match(rp.reshape(64, 64, 16), pcf, ctr.float())
    gidx, gxy = net.fine_match(gp.reshape(64, 64, 16), pcf.cuda(), ctr.float().cuda())
    assert torch.equal(gidx.cpu(), ridx) and torch.equal(gxy.cpu(), rxy)
    with pytest.raises(AssertionError):
        net.extract_patch(fmap.cuda(), torch.tensor([[1.0], [40.0]]).cuda())


def test_cpu_tensors_rejected():
    ops = _ops()
    with pytest.raises(RuntimeError):
        ops.gemm(torch.zeros(4, 4), torch.zeros(4, 4))
