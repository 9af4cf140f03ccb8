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
        l4 = self.decoder4(self._up_cat(f5, up[3], f4, f), f)
        l3 = self.decoder3(self._up_cat(l4, up[2], f3, f), f)
        l2 = self.decoder2(self._up_cat(l3, up[1], f2, f), f)
        if taps is not None:
            taps.update(encoder1_2=f1, encoder2_3=f2, encoder3_3=f3, encoder4_3=f4, encoder5_3=f5,
                        decoder4=l4, decoder3=l3, decoder2=l2)
        return [l2, l3, l4, f5]
