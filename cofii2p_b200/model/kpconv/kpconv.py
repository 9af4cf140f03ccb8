"""Rigid KPConv on the B200 kernels (reference model/kpconv/kpconv.py:10-122)."""
import math

import torch
import torch.nn as nn

from ... import autograd as ad
from ... import ops
from .kernel_points import load_kernels


class KPConv(nn.Module):
    """Same parameters/buffers as the reference: `weights` [K,Cin,Cout], `bias` [Cout] or None,
    buffer `kernel_points` [K,3]."""

    def __init__(self, in_channels, out_channels, kernel_size, radius, sigma, bias=False, dimension=3,
                 inf=1e6, eps=1e-9):
        super().__init__()
        self.kernel_size, self.in_channels, self.out_channels = kernel_size, in_channels, out_channels
        self.radius, self.sigma, self.dimension, self.inf, self.eps = radius, sigma, dimension, inf, eps
        self.weights = nn.Parameter(torch.zeros(kernel_size, in_channels, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()
        self.register_buffer("kernel_points",
                             torch.from_numpy(load_kernels(radius, kernel_size, dimension=dimension, fixed="center")).float())
        self._wt = None
        self._wt_version = None
        self._reach = None
        self._wt16 = None

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weights, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.weights)
            bound = 1 / math.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)

    def packed_weight(self):
        """[Cout, K*Cin]: the K-major operand of the weight-apply GEMM (cached per parameter version)."""
        v = (self.weights._version, self.weights.data_ptr(), ops.weights_epoch())
        if self._wt is None or self._wt_version != v:
            with torch.no_grad():
                self._wt = self.weights.detach().reshape(-1, self.out_channels).t().contiguous()
            self._wt_version = v
        return self._wt

    def packed_weight_f16(self):
        """fp16 copy of the K-major weight operand (tf32 engine), cached per parameter version."""
        wt = self.packed_weight()
        if self._wt16 is None or self._wt16[0] is not wt:
            self._wt16 = (wt, wt.to(torch.float16))
        return self._wt16[1]

    def kp_reach(self) -> float:
        """max_k |kernel_points[k]| (host scalar, cached): lets the kernel cull neighbours that no kernel point reaches."""
        v = (self.kernel_points._version, self.kernel_points.data_ptr())
        if self._reach is None or self._reach[0] != v:
            self._reach = (v, float(self.kernel_points.detach().norm(dim=1).max().item()))
        return self._reach[1]

    def forward(self, s_feats, q_points, s_points, neighbor_indices, frames: int = 1, want_stats: bool = False):
        """s_feats [B*N,Cin], q_points [B*M,3], s_points [B*N,3], neighbor_indices [B*M,H] -> [B*M,Cout]."""
        with ops.group("kpconv"):
            return self._forward(s_feats, q_points, s_points, neighbor_indices, frames, want_stats)

    def _forward(self, s_feats, q_points, s_points, neighbor_indices, frames, want_stats):
        if ad.active(self):  # training: differentiable kernels (fp32 aggregate, engine-selected GEMMs)
            if want_stats and ops.colstats_ok(q_points.shape[0], frames, self.out_channels):
                return ad.kpconv(s_feats, self.weights, self.bias, q_points, s_points, neighbor_indices, self.kernel_points,
                                 self.sigma, frames, self.kp_reach(), want_stats=True)
            out = ad.kpconv(s_feats, self.weights, self.bias, q_points, s_points, neighbor_indices, self.kernel_points,
                            self.sigma, frames, self.kp_reach())
            return (out, None) if want_stats else out
        packed = ops.pack_points(s_points, s_feats)
        if ops.engine_id() == ops.ENGINE_TF32 and (self.in_channels * self.kernel_size) % 8 == 0 and self.out_channels >= 16:
            # tf32 engine: fp16 aggregate (same 11-bit operand precision as tf32), half the HBM round trip
            agg, cnt = ops.kpconv_aggregate_f16(s_feats, packed, q_points, neighbor_indices, self.kernel_points,
                                                self.sigma, frames, self.kp_reach())
            if want_stats and ops.colstats_ok(q_points.shape[0], frames, self.out_channels):
                return ops.gemm_f16_colstats(agg, self.packed_weight_f16(), bias=self.bias, rowdiv=cnt)
            out = ops.gemm_f16(agg, self.packed_weight_f16(), bias=self.bias, rowdiv=cnt)
            return (out, None) if want_stats else out
        agg, cnt = ops.kpconv_aggregate(s_feats, packed, q_points, neighbor_indices, self.kernel_points, self.sigma,
                                        frames, self.kp_reach())
        if want_stats and ops.colstats_ok(q_points.shape[0], frames, self.out_channels):
            return ops.gemm_colstats(agg, self.packed_weight(), bias=self.bias, rowdiv=cnt, const_w=True)
        out = ops.gemm(agg, self.packed_weight(), bias=self.bias, rowdiv=cnt, const_w=True)
        return (out, None) if want_stats else out
