"""TEST INFRASTRUCTURE ONLY -- freezes outputs of the *real reference* into tests/golden/*.npz.

Runs in the build container only (needs /root/reference, see oracle/ref_shim.py):
    python -m oracle.make_golden
For each (frame seed, cloud size) the unmodified reference `CoFiI2P.forward` is run on CPU in `val` and `test`
modes with the seeded state_dict of `cofii2p_b200.weights.seeded_state_dict(model, seed=0)` loaded through
`load_state_dict(strict=True)`; the 8 outputs are stored together with strided samples of per-module
intermediates (hooks on the reference modules).  The GPU box rebuilds the identical inputs and weights from
the seeds and compares the CUDA path against these files (tests/test_golden_gpu.py); the CPU suite checks
`oracle/restate.py` against them (tests/test_oracle.py).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cofii2p_b200.frames import make_frame  # noqa: E402
from cofii2p_b200.model.network import CoFiI2P  # noqa: E402  (construction only: CPU, no kernels run)
from cofii2p_b200.options import Options_KITTI  # noqa: E402
from cofii2p_b200.weights import seeded_state_dict  # noqa: E402
from oracle.ref_shim import build_reference_model  # noqa: E402

CASES = [(0, 4096), (1, 4096), (0, 20480)]  # (frame seed, num_pc)
TAPS = ["pc_encoder.encoder1_1", "pc_encoder.encoder1_2", "pc_encoder.encoder2_3", "pc_encoder.encoder3_3",
        "pc_encoder.encoder4_3", "pc_encoder.encoder5_3", "pc_encoder.decoder4", "pc_encoder.decoder3",
        "pc_encoder.decoder2", "pc_feature_layer", "img_upsample_1", "img_upsample_2",
        "img_encoder.backbone.layer1", "img_encoder.backbone.layer2"]


def sample(t: torch.Tensor, n: int = 4096) -> np.ndarray:
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n].numpy().copy()


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    net, _ = build_reference_model(0)
    sd = seeded_state_dict(CoFiI2P(Options_KITTI()), 0)
    net.load_state_dict(sd, strict=True)
    net.eval()
    mods = dict(net.named_modules())
    for seed, num_pc in CASES:
        f = make_frame(seed, num_pc=num_pc, cache_dir="/tmp/cofi_frames")
        args = (f["pc_data_dict"], f["img"], f["fine_center_kpt_coors"], f["fine_xy"], f["fine_pc_inline_index"])
        rec = {}
        taps = {}
        hooks = [mods[name].register_forward_hook(lambda m, i, o, name=name: taps.__setitem__(name, o)) for name in TAPS]
        with torch.no_grad():
            val = net(*args, "val")
        for h in hooks:
            h.remove()
        for name, o in taps.items():
            rec["tap/" + name] = sample(o)
            rec["tapshape/" + name] = np.array(o.shape)
        with torch.no_grad():
            test = net(*args, "test")
        names = ["img_feature_norm", "pc_feature_norm", "coarse_img_score", "coarse_pc_score",
                 "fine_img_feature_patch", "fine_pc_inline_feature", "fine_center_xy", "coarse_pc_points"]
        for i, nm in enumerate(names):
            if val[i] is not None:
                rec["val/" + nm] = val[i].numpy()
            if test[i] is not None:
                rec["test/" + nm] = test[i].numpy()
        for i in range(4):
            assert torch.equal(val[i], test[i])
            del rec["test/" + names[i]]
        path = os.path.join(out_dir, f"frame_s{seed}_n{num_pc}.npz")
        np.savez_compressed(path, **rec)
        print(path, os.path.getsize(path) // 1024, "KiB; test-mode matches:", test[6].shape[1],
              "score>=0.9:", int((val[3] >= 0.9).sum()))


if __name__ == "__main__":
    main()
