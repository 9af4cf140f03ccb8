"""reference model/kpconv/functional.py:5-21,53-66 on the B200 kernels."""
import torch

from ... import autograd as ad
from ... import ops


def _grad(x):
    return torch.is_grad_enabled() and x.requires_grad


def nearest_upsample(x, upsample_indices, frames: int = 1, out=None):
    """Only column 0 of the [n2, max_num] table is read (the tables are distance-ordered)."""
    if _grad(x):
        g = ad.gather(x, upsample_indices, upsample_indices.shape[1], frames, upsample_indices.shape[0])
        if out is not None:
            raise RuntimeError("nearest_upsample: `out` is an inference-only fast path")
        return g
    return ops.gather_rows(x, upsample_indices, idx_stride=upsample_indices.shape[1], frames=frames, out=out,
                           rows_out=upsample_indices.shape[0])


def maxpool(x, neighbor_indices, frames: int = 1):
    if _grad(x):
        return ad.maxpool_rows(x, neighbor_indices, frames)
    with ops.group("pc_maxpool"):
        if ops.engine_id() == ops.ENGINE_TF32 and x.shape[1] % 8 == 0:
            # tf32 engine: gather an fp16 copy (monotonic rounding => exact max of the rounded rows), half the L2 bytes
            return ops.maxpool_rows_f16(ops.cast_f16(x), neighbor_indices, frames)
        return ops.maxpool_rows(x, neighbor_indices, frames)
