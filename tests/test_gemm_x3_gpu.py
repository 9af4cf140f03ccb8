"""Persistent 3xTF32 engine with pre-split weights (csrc/gemm_x3.cu, COFI_GEMM_TF32X3S: A split in registers into
tensor memory, W split once by cofi_split_tf32) against fp64 torch.  Tolerance: fp32-grade -- what is left after the
operand split is the tensor core's truncating fp32 accumulator, 2e-6 + 1.2e-8 * K of the output scale (tools/err_probe.py)."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu


def tol(k):
    return 2e-6 + 1.2e-8 * k


def test_split_tf32_planes():
    from cofii2p_b200 import ops
    w = torch.randn(70, 52, device="cuda") * torch.logspace(-6, 3, 52, device="cuda")
    s = ops.split_tf32(w)
    assert s.shape == (2, 70, 52)
    bits = s.view(torch.int32)
    assert int((bits & 0x1FFF).abs().max()) == 0            # both planes are exact tf32 values
    err = (w.double() - s[0].double() - s[1].double()).abs() / w.double().abs().clamp_min(1e-30)
    assert float(err.max()) < 2.0 ** -21


@pytest.mark.parametrize("m,n,k", [(1000, 64, 60), (513, 32, 480), (2000, 128, 32), (300, 1024, 3072), (4096, 2048, 1024),
                                   (77, 200, 256), (128, 48, 36), (20480, 32, 64), (1280, 256, 128), (163840, 128, 32),
                                   (40960, 512, 128), (20480, 128, 128), (19000, 96, 4)])
def test_gemm_x3s(m, n, k):
    from cofii2p_b200 import ops
    g = torch.Generator().manual_seed(m + n + k)
    a = torch.randn((m, k), generator=g)
    w = torch.randn((n, k), generator=g) / math.sqrt(k)
    b = torch.randn((n,), generator=g)
    rd = torch.randint(1, 60, (m,), generator=g).float()
    ref = F.leaky_relu(F.linear(a.double(), w.double()).float() / rd[:, None] + b, 0.1)
    ops.set_engine("tf32x3")
    try:
        ac, wc = a.cuda(), w.cuda()
        got = ops.gemm(ac, wc, bias=b.cuda(), rowdiv=rd.cuda(), act=ops.ACT_LRELU, const_w=True)
        assert rel_err(got, ref) < tol(k), rel_err(got, ref)
        lin = F.linear(a.double(), w.double()).float()
        base = torch.randn((m, n), generator=g)
        got2 = ops.gemm(ac, wc, out=base.clone().cuda(), accumulate=True, const_w=True)
        assert rel_err(got2, base + lin) < tol(k)
        # strided A (view into a wider buffer) and strided output; then an output whose rows are not 16-byte aligned
        wide = torch.randn((m, k + 36), generator=g).cuda()
        outw = torch.zeros((m, n + 8), device="cuda")
        ops.gemm(wide[:, 4:4 + k], wc, out=outw[:, 4:4 + n], const_w=True)
        assert rel_err(outw[:, 4:4 + n], F.linear(wide[:, 4:4 + k].cpu().double(), w.double()).float()) < tol(k)
        assert float(outw[:, :4].abs().max()) == 0 and float(outw[:, 4 + n:].abs().max()) == 0
        outo = torch.zeros((m, n + 3), device="cuda")
        ops.gemm(ac, wc, out=outo[:, 1:1 + n], const_w=True)
        assert rel_err(outo[:, 1:1 + n], lin) < tol(k)
        assert float(outo[:, :1].abs().max()) == 0 and float(outo[:, 1 + n:].abs().max()) == 0
        # same numbers as the one-tile-per-CTA 3xTF32 kernel that splits both operands in shared memory
        old = ops.gemm(ac, wc, const_w=False)
        new = ops.gemm(ac, wc, const_w=True)
        assert rel_err(new, old.cpu()) < tol(k)
    finally:
        ops.set_engine("fp32")


@pytest.mark.parametrize("m,n,k", [(1280, 128, 256), (1000, 64, 128), (333, 32, 64), (20480, 128, 128), (40000, 96, 256)])
def test_gemm_ln_x3s(m, n, k):
    from cofii2p_b200 import ops
    g = torch.Generator().manual_seed(m + n + k)
    a = torch.randn((m, k), generator=g)
    w = torch.nn.Parameter(torch.randn((n, k), generator=g) / math.sqrt(k))
    gamma, beta = torch.randn((n,), generator=g), torch.randn((n,), generator=g)
    res = torch.randn((m, n), generator=g)
    ref = F.relu(F.layer_norm(F.linear(a.double(), w.detach().double()), (n,), gamma.double(), beta.double(), 1e-5)).float() + res
    ops.set_engine("tf32x3")
    try:
        with torch.no_grad():
            wc = torch.nn.Parameter(w.detach().cuda())
            got = ops.gemm_ln(a.cuda(), wc, gamma.cuda(), beta.cuda(), 1e-5, act=ops.ACT_RELU, residual=res.cuda())
    finally:
        ops.set_engine("fp32")
    assert rel_err(got, ref) < 1e-4, rel_err(got, ref)


@pytest.mark.parametrize("rows,frames,k,n,groups", [(2560, 2, 64, 128, 32), (1280, 1, 480, 32, 32), (1024, 8, 128, 256, 32),
                                                    (20480, 8, 32, 128, 32)])
def test_gemm_colstats_x3s(rows, frames, k, n, groups):
    from cofii2p_b200 import ops
    g = torch.Generator().manual_seed(rows + k + n)
    a = torch.randn((rows * frames, k), generator=g)
    w = torch.randn((n, k), generator=g) / math.sqrt(k)
    b = torch.randn((n,), generator=g)
    gamma, beta = torch.randn((n,), generator=g), torch.randn((n,), generator=g)
    ops.set_engine("tf32x3")
    try:
        y, st = ops.gemm_colstats(a.cuda(), w.cuda(), bias=b.cuda(), const_w=True)
        y2 = ops.gemm(a.cuda(), w.cuda(), bias=b.cuda(), const_w=True)
        assert torch.equal(y, y2)
        assert rel_err(st[..., 0].sum(0), y.double().sum(0).float()) < 1e-4
        assert rel_err(st[..., 1].sum(0), (y.double() ** 2).sum(0).float()) < 1e-4
        got = ops.norm_rows_pre(y, st, frames, groups, gamma.cuda(), beta.cuda(), 1e-5, act=ops.ACT_LRELU)
        ref2 = ops.norm_rows(y2, frames, groups, gamma.cuda(), beta.cuda(), 1e-5, act=ops.ACT_LRELU)
        assert rel_err(got, ref2.cpu()) < 1e-5
    finally:
        ops.set_engine("fp32")
    assert rel_err(y, F.linear(a.double(), w.double(), b.double()).float()) < tol(k)


def test_split_cache_follows_the_weights_epoch():
    from cofii2p_b200 import ops
    w = torch.nn.Parameter(torch.randn(64, 64, device="cuda"))
    a = torch.randn(256, 64, device="cuda")
    ops.set_engine("tf32x3")
    try:
        with torch.no_grad():
            y0 = ops.gemm(a, w)
            w.data.mul_(2.0)            # raw update, as the fused Adam kernel does
            ops.bump_weights_epoch()
            y1 = ops.gemm(a, w)
    finally:
        ops.set_engine("fp32")
    assert rel_err(y1, (2.0 * y0).cpu()) < 1e-6
