"""KPConv building blocks (reference model/kpconv/modules.py:32-240). nn.Linear / nn.GroupNorm instances are
parameter containers only (identical state_dict keys); the arithmetic runs in libcofi_b200.so."""
import torch
import torch.nn as nn

from ... import autograd as ad
from ... import ops
from .functional import maxpool, nearest_upsample
from .kpconv import KPConv


class MaxPool(nn.Module):
    @staticmethod
    def forward(s_feats, neighbor_indices, frames: int = 1):
        return maxpool(s_feats, neighbor_indices, frames)


class GroupNorm(nn.Module):
    def __init__(self, num_groups, num_channels):
        super().__init__()
        self.num_groups, self.num_channels = num_groups, num_channels
        self.norm = nn.GroupNorm(num_groups, num_channels)

    def forward(self, x, frames: int = 1, act: int = ops.ACT_NONE, residual=None, tile_stats=None):
        if ad.active(self):
            if tile_stats is not None:
                return ad.norm_rows_pre(x, tile_stats, frames, self.num_groups, self.norm.weight, self.norm.bias, self.norm.eps,
                                        residual, act)
            return ad.norm_rows(x, frames, self.num_groups, self.norm.weight, self.norm.bias, self.norm.eps, residual, act)
        if tile_stats is not None:  # statistics came out of the producing GEMM's epilogue
            return ops.norm_rows_pre(x, tile_stats, frames, self.num_groups, self.norm.weight, self.norm.bias, self.norm.eps,
                                     residual=residual, act=act)
        return ops.norm_rows(x, frames, self.num_groups, self.norm.weight, self.norm.bias, self.norm.eps,
                             residual=residual, act=act)


def get_norm(norm, channels, num_groups=None):
    if norm == "gn":
        assert num_groups > 1, "number of groups must be positive number!"
        return GroupNorm(num_groups, channels)
    raise ValueError("cofii2p_b200 implements the reference's default point-branch normalisation ('gn') only")


class UnaryBlock(nn.Module):
    def __init__(self, in_channels, out_channels, norm, group_norm, has_relu=True, bias=True, layer_norm=False):
        super().__init__()
        self.in_channels, self.out_channels, self.group_norm = in_channels, out_channels, group_norm
        self.mlp = nn.Linear(in_channels, out_channels, bias=bias)
        self.norm = get_norm(norm, out_channels, num_groups=group_norm)
        self.leaky_relu = nn.LeakyReLU(0.1) if has_relu else None

    def forward(self, x, frames: int = 1, residual=None, final_act: int = None):
        with ops.group("pc_unary"):
            return self._forward(x, frames, residual, final_act)

    def _forward(self, x, frames, residual, final_act):
        act = ops.ACT_LRELU if self.leaky_relu is not None else ops.ACT_NONE
        if final_act is not None:
            act = final_act
        if ad.active(self):
            if ops.colstats_ok(x.shape[0], frames, self.out_channels):
                y, st = ad.linear_stats(x, self.mlp.weight, self.mlp.bias)
                return self.norm(y, frames, act=act, residual=residual, tile_stats=st)
            return self.norm(ad.linear(x, self.mlp.weight, self.mlp.bias), frames, act=act, residual=residual)
        if ops.colstats_ok(x.shape[0], frames, self.out_channels):
            y, st = ops.gemm_colstats(x, self.mlp.weight, bias=self.mlp.bias)
            return self.norm(y, frames, act=act, residual=residual, tile_stats=st)
        y = ops.gemm(x, self.mlp.weight, bias=self.mlp.bias)
        return self.norm(y, frames, act=act, residual=residual)


class LastUnaryBlock(nn.Module):
    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.mlp = nn.Linear(in_channels, out_channels, bias=bias)

    def forward(self, x, frames: int = 1):
        with ops.group("pc_unary"):
            if ad.active(self):
                return ad.linear(x, self.mlp.weight, self.mlp.bias)
            return ops.gemm(x, self.mlp.weight, bias=self.mlp.bias)


class ConvBlock(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, radius, sigma, norm, group_norm, negative_slope=0.1,
                 bias=True, layer_norm=False):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.KPConv = KPConv(in_channels, out_channels, kernel_size, radius, sigma, bias=bias)
        self.norm = get_norm(norm, out_channels, group_norm)
        self.leaky_relu = nn.LeakyReLU(negative_slope=negative_slope)

    def forward(self, s_feats, q_points, s_points, neighbor_indices, frames: int = 1):
        x, st = self.KPConv(s_feats, q_points, s_points, neighbor_indices, frames, want_stats=True)
        return self.norm(x, frames, act=ops.ACT_LRELU, tile_stats=st)


class ResidualBlock(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, radius, sigma, norm, group_norm, strided=False,
                 bias=True, layer_norm=False):
        super().__init__()
        self.in_channels, self.out_channels, self.strided = in_channels, out_channels, strided
        mid = out_channels // 4
        self.unary1 = (UnaryBlock(in_channels, mid, norm, group_norm, bias=bias, layer_norm=layer_norm)
                       if in_channels != mid else nn.Identity())
        self.KPConv = KPConv(mid, mid, kernel_size, radius, sigma, bias=bias)
        self.norm_conv = get_norm(norm, mid, group_norm)
        self.unary2 = UnaryBlock(mid, out_channels, norm, group_norm, has_relu=False, bias=bias, layer_norm=layer_norm)
        self.unary_shortcut = (UnaryBlock(in_channels, out_channels, norm, group_norm, has_relu=False, bias=bias,
                                          layer_norm=layer_norm) if in_channels != out_channels else nn.Identity())
        self.leaky_relu = nn.LeakyReLU(0.1)

    def forward(self, s_feats, q_points, s_points, neighbor_indices, frames: int = 1):
        x = self.unary1(s_feats, frames) if isinstance(self.unary1, UnaryBlock) else s_feats
        x, st = self.KPConv(x, q_points, s_points, neighbor_indices, frames, want_stats=True)
        x = self.norm_conv(x, frames, act=ops.ACT_LRELU, tile_stats=st)
        shortcut = maxpool(s_feats, neighbor_indices, frames) if self.strided else s_feats
        if isinstance(self.unary_shortcut, UnaryBlock):
            shortcut = self.unary_shortcut(shortcut, frames)
        # x = LeakyReLU(GN(Linear(x)) + shortcut): the add and the activation ride in the norm-apply kernel
        return self.unary2(x, frames, residual=shortcut, final_act=ops.ACT_LRELU)
