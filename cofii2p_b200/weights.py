"""Deterministic, reference-independent weights for tests and benchmarks.

No checkpoint is reachable (the reference's weights live on web drives), so parity and throughput are
measured on a seeded random `state_dict` that is a pure function of (key order, shapes, seed): the same
tensors are rebuilt bit for bit in the build container (where the real reference consumes them through
`load_state_dict(strict=True)` to produce the goldens) and on the GPU box.  Scales follow the usual fan-in
rules so activations stay O(1); normalisation affines and BatchNorm running statistics are non-trivial on
purpose (a stricter test than the default 1/0).  The last conv of the point score head is scaled up so that
a healthy fraction of super-points passes the 0.9 test-mode threshold."""
from collections import OrderedDict

import math
import numpy as np
import torch

from .model.kpconv.kernel_points import _disposition


def seeded_state_dict(model: torch.nn.Module, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    ref = model.state_dict()
    out = OrderedDict()
    for i, (k, v) in enumerate(ref.items()):
        g = torch.Generator().manual_seed(seed * 100003 + i * 7919 + 17)
        shape = tuple(v.shape)
        if k.endswith("num_batches_tracked"):
            t = torch.zeros(shape, dtype=v.dtype)
        elif k.endswith("kernel_points"):
            # radius is recoverable from the block: kernel points scale with the layer radius
            radius = float(v.norm(dim=1)[1:].mean() / 0.66) if v.shape[0] > 1 else 1.0
            radius = round(radius / 0.425) * 0.425 if radius > 0.2 else 0.425
            base = torch.from_numpy(_disposition(shape[0])).to(torch.float32)
            th = float(torch.rand((), generator=g)) * 2 * math.pi
            c, s = math.cos(th), math.sin(th)
            R = torch.tensor([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]], dtype=torch.float32)
            t = (radius * (base + 0.01 * torch.randn(shape, generator=g))) @ R
        elif k.endswith("running_mean"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif k.endswith("running_var"):
            t = 0.5 + torch.rand(shape, generator=g)
        elif v.dim() == 1:
            if k.endswith("weight"):
                t = 1.0 + 0.1 * torch.randn(shape, generator=g)
            else:
                t = 0.1 * torch.randn(shape, generator=g)
        else:
            if k.endswith("KPConv.weights"):
                fan_in = shape[1] * 2.0  # ~2 kernel points influence a neighbour on average
            else:
                fan_in = int(np.prod(shape[1:]))
            bound = math.sqrt(3.0 / max(fan_in, 1))
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
            if k == "pc_score_layer.6.weight" or k == "img_score_layer.6.weight":
                t = t * 6.0
        out[k] = t.to(v.dtype)
    return out
