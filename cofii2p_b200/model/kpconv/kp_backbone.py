"""KPConvFPN: 5-stage encoder / 3-stage decoder (reference model/kpconv/kp_backbone.py:8-128)."""
import torch
import torch.nn as nn

from ... import ops
from .modules import ConvBlock, ResidualBlock, UnaryBlock, LastUnaryBlock
from .functional import nearest_upsample


class KPConvFPN(nn.Module):
    def __init__(self, input_dim, output_dim, init_dim, kernel_size, init_radius, init_sigma, norm, group_norm):
        super().__init__()
        d, k, r, s = init_dim, kernel_size, init_radius, init_sigma
        a = (norm, group_norm)
        self.encoder1_1 = ConvBlock(input_dim, d, k, r, s, *a)
        self.encoder1_2 = ResidualBlock(d, d * 2, k, r, s, *a)
        self.encoder2_1 = ResidualBlock(d * 2, d * 2, k, r, s, *a, strided=True)
        self.encoder2_2 = ResidualBlock(d * 2, d * 4, k, r * 2, s * 2, *a)
        self.encoder2_3 = ResidualBlock(d * 4, d * 4, k, r * 2, s * 2, *a)
        self.encoder3_1 = ResidualBlock(d * 4, d * 4, k, r * 2, s * 2, *a, strided=True)
        self.encoder3_2 = ResidualBlock(d * 4, d * 8, k, r * 4, s * 4, *a)
        self.encoder3_3 = ResidualBlock(d * 8, d * 8, k, r * 4, s * 4, *a)
        self.encoder4_1 = ResidualBlock(d * 8, d * 8, k, r * 4, s * 4, *a, strided=True)
        self.encoder4_2 = ResidualBlock(d * 8, d * 16, k, r * 8, s * 8, *a)
        self.encoder4_3 = ResidualBlock(d * 16, d * 16, k, r * 8, s * 8, *a)
        self.encoder5_1 = ResidualBlock(d * 16, d * 16, k, r * 8, s * 8, *a, strided=True)
        self.encoder5_2 = ResidualBlock(d * 16, d * 32, k, r * 16, s * 16, *a)
        self.encoder5_3 = ResidualBlock(d * 32, d * 32, k, r * 16, s * 16, *a)
        self.decoder4 = UnaryBlock(d * 48, d * 16, *a)
        self.decoder3 = UnaryBlock(d * 24, d * 8, *a)
        self.decoder2 = LastUnaryBlock(d * 12, output_dim)

    def _up_cat(self, coarse, up_table, skip, frames):
        """torch.cat([nearest_upsample(coarse, up), skip], 1) written straight into one buffer."""
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
