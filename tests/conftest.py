import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FRAME_CACHE = os.environ.get("COFI_FRAME_CACHE", "/tmp/cofi_frames")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def seeded_sd():
    """Reference-independent seeded state_dict (cofii2p_b200.weights), CPU tensors."""
    from cofii2p_b200.model.network import CoFiI2P
    from cofii2p_b200.options import Options_KITTI
    from cofii2p_b200.weights import seeded_state_dict
    return seeded_state_dict(CoFiI2P(Options_KITTI()), 0)


@pytest.fixture(scope="session")
def cuda_model(seeded_sd):
    from cofii2p_b200.model.network import CoFiI2P
    from cofii2p_b200.options import Options_KITTI
    m = CoFiI2P(Options_KITTI())
    m.load_state_dict(seeded_sd, strict=True)
    return m.cuda().eval()


def get_frame(seed, num_pc):
    from cofii2p_b200.frames import make_frame
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    return make_frame(seed, num_pc=num_pc, cache_dir=FRAME_CACHE, device=dev)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| relative to the scale of the reference tensor b."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
