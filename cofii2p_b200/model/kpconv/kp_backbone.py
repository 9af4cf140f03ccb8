"""KPConvFPN: 5-stage encoder / 3-stage decoder (reference model/kpconv/kp_backbone.py:8-128)."""
import torch
import torch.nn as nn

from ... import autograd as ad
from ... import ops
from .modules import ConvBlock, ResidualBlock, UnaryBlock, LastUnaryBlock
from .functional import nearest_upsample


# (block name, in/out channels as multiples of init_dim, radius/sigma multiplier, strided); strided blocks keep the
# finer stage's radius (reference kp_backbone.py:11-73).  encoder1_1 is the only plain ConvBlock.
_ENCODER = (("encoder1_2", 1, 2, 1, False),
            ("encoder2_1", 2, 2, 1, True), ("encoder2_2", 2, 4, 2, False), ("encoder2_3", 4, 4, 2, False),
            ("encoder3_1", 4, 4, 2, True), ("encoder3_2", 4, 8, 4, False), ("encoder3_3", 8, 8, 4, False),
            ("encoder4_1", 8, 8, 4, True), ("encoder4_2", 8, 16, 8, False), ("encoder4_3", 16, 16, 8, False),
            ("encoder5_1", 16, 16, 8, True), ("encoder5_2", 16, 32, 16, False), ("encoder5_3", 32, 32, 16, False))


class KPConvFPN(nn.Module):
    def __init__(self, input_dim, output_dim, init_dim, kernel_size, init_radius, init_sigma, norm, group_norm):
        super().__init__()
        self.encoder1_1 = ConvBlock(input_dim, init_dim, kernel_size, init_radius, init_sigma, norm, group_norm)
        for name, cin, cout, scale, strided in _ENCODER:  # registration order == the reference's state_dict order
            setattr(self, name, ResidualBlock(init_dim * cin, init_dim * cout, kernel_size, init_radius * scale,
                                              init_sigma * scale, norm, group_norm, strided=strided))
        self.decoder4 = UnaryBlock(init_dim * (32 + 16), init_dim * 16, norm, group_norm)
        self.decoder3 = UnaryBlock(init_dim * (16 + 8), init_dim * 8, norm, group_norm)
        self.decoder2 = LastUnaryBlock(init_dim * (8 + 4), output_dim)

    def _up_cat(self, coarse, up_table, skip, frames):
        """torch.cat([nearest_upsample(coarse, up), skip], 1) written straight into one buffer."""
        if ad.active(self):
            return torch.cat([nearest_upsample(coarse, up_table, frames), skip], 1)
        n, c1, c2 = skip.shape[0], coarse.shape[1], skip.shape[1]
        buf = torch.empty((n, c1 + c2), dtype=torch.float32, device=skip.device)
        nearest_upsample(coarse, up_table, frames, out=buf[:, :c1])
        ops.gather_rows(skip, None, frames=frames, out=buf[:, c1:])
        return buf

    def _decode(self, block, coarse, up_table, skip, frames):
        """block(cat([nearest_upsample(coarse), skip])) (reference kp_backbone.py:100-118).  Inference: a row gather commutes
        with a row-wise linear map, W [up(x_c) | x_f] = up(W_c x_c) + W_f x_f, so the coarse half of the Linear runs at the
        coarse resolution (half the rows: a third fewer flops), its result is up-sampled straight into the output buffer and
        the fine half accumulates onto it -- the concatenated [rows, c1 + c2] buffer never exists."""
        if ad.active(self):
            return block(self._up_cat(coarse, up_table, skip, frames), frames)
        with ops.group("pc_unary"):
            W, c1 = block.mlp.weight, coarse.shape[1]
            yc = ops.gemm(coarse, W[:, :c1])
            y = nearest_upsample(yc, up_table, frames)
            if isinstance(block, UnaryBlock) and ops.colstats_ok(skip.shape[0], frames, block.out_channels):
                y, st = ops.gemm_colstats(skip, W[:, c1:], bias=block.mlp.bias, out=y, accumulate=True)
                return block.norm(y, frames, act=ops.ACT_LRELU if block.leaky_relu is not None else ops.ACT_NONE, tile_stats=st)
            y = ops.gemm(skip, W[:, c1:], bias=block.mlp.bias, out=y, accumulate=True)
            if isinstance(block, UnaryBlock):
                return block.norm(y, frames, act=ops.ACT_LRELU if block.leaky_relu is not None else ops.ACT_NONE)
            return y

    def forward(self, data_dict, frames: int = 1, taps=None):
        feats = data_dict["feats"]
        p, nb = data_dict["points"], data_dict["neighbors"]
        sub, up = data_dict["subsampling"], data_dict["upsampling"]
        f = frames
        f1 = self.encoder1_1(feats, p[0], p[0], nb[0], f)
        f1 = self.encoder1_2(f1, p[0], p[0], nb[0], f)
        f2 = self.encoder2_1(f1, p[1], p[0], sub[0], f)
        f2 = self.encoder2_2(f2, p[1], p[1], nb[1], f)
        f2 = self.encoder2_3(f2, p[1], p[1], nb[1], f)
        f3 = self.encoder3_1(f2, p[2], p[1], sub[1], f)
        f3 = self.encoder3_2(f3, p[2], p[2], nb[2], f)
        f3 = self.encoder3_3(f3, p[2], p[2], nb[2], f)
        f4 = self.encoder4_1(f3, p[3], p[2], sub[2], f)
        f4 = self.encoder4_2(f4, p[3], p[3], nb[3], f)
        f4 = self.encoder4_3(f4, p[3], p[3], nb[3], f)
        f5 = self.encoder5_1(f4, p[4], p[3], sub[3], f)
        f5 = self.encoder5_2(f5, p[4], p[4], nb[4], f)
        f5 = self.encoder5_3(f5, p[4], p[4], nb[4], f)
        l4 = self._decode(self.decoder4, f5, up[3], f4, f)
        l3 = self._decode(self.decoder3, l4, up[2], f3, f)
        l2 = self._decode(self.decoder2, l3, up[1], f2, f)
        if taps is not None:
            taps.update(encoder1_2=f1, encoder2_3=f2, encoder3_3=f3, encoder4_3=f4, encoder5_3=f5,
                        decoder4=l4, decoder3=l3, decoder2=l2)
        return [l2, l3, l4, f5]
