"""Positional encodings (reference model/transformer/position_encoding.py:7-70)."""
import math

import torch
import torch.nn as nn

from ... import ops


class PositionEmbeddingCoordsSine(nn.Module):
    def __init__(self, n_dim: int = 1, d_model: int = 64, temperature=10000, scale=None):
        super().__init__()
        self.n_dim = n_dim
        self.num_pos_feats = d_model // n_dim // 2 * 2
        self.temperature = temperature
        self.padding = d_model - self.num_pos_feats * self.n_dim
        self.d_model = d_model
        self.scale = (1.0 if scale is None else scale) * 2 * math.pi
        self._dim_t = {}

    def dim_t(self, device):
        """temperature ** (2*(i//2)/npf), built with the same torch ops as the reference (:39-40) so the divisor
        is bit-identical."""
        key = str(device)
        if key not in self._dim_t:
            t = torch.arange(self.num_pos_feats, dtype=torch.float32)
            t = self.temperature ** (2 * torch.div(t, 2, rounding_mode="trunc") / self.num_pos_feats)
            self._dim_t[key] = t.to(device)
        return self._dim_t[key]

    def forward(self, xyz: torch.Tensor) -> torch.Tensor:
        """xyz (*, n_dim) -> (*, d_model)."""
        assert xyz.shape[-1] == self.n_dim
        lead = xyz.shape[:-1]
        out = ops.posenc_sine(xyz.reshape(-1, self.n_dim).to(torch.float32), self.d_model, self.dim_t(xyz.device))
        return out.view(*lead, self.d_model)


class PositionEmbeddingLearned(nn.Module):
    """Instantiated by the reference (model/network.py:36-37) but never called: parameters only."""

    def __init__(self, n_dim: int = 1, d_model: int = 64):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(n_dim, 32), nn.ReLU(), nn.Linear(32, 64), nn.ReLU(), nn.Linear(64, 128),
                                 nn.ReLU(), nn.Linear(128, 256), nn.ReLU(), nn.Linear(256, d_model))

    def forward(self, xyz):
        raise RuntimeError("PositionEmbeddingLearned is dead code on the hot path (reference network.py:36-37)")
