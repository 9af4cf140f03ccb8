"""Full softmax attention (reference model/transformer/linear_attention.py:50-79) as one fused kernel."""
import torch.nn as nn

from ... import ops


class FullAttention(nn.Module):
    def __init__(self, use_dropout=False, attention_dropout=0.1):
        super().__init__()
        assert not use_dropout, "the reference never enables attention dropout on this path"
        self.use_dropout = use_dropout

    def forward(self, queries, keys, values, frames: int = 1, heads: int = 4):
        """queries [B*L, H*D], keys/values [B*S, H*D] -> [B*L, H*D]; softmax temperature 1/sqrt(D)."""
        D = queries.shape[1] // heads
        return ops.attention(queries, keys, values, frames, heads, 1.0 / D ** 0.5)
