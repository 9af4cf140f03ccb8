"""Per-parameter gradient comparison: CUDA training path vs torch autograd through the CPU oracle (debug aid)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from conftest import get_frame
    from cofii2p_b200 import ops
    from cofii2p_b200.frames import frame_to, stack_frames
    from cofii2p_b200.model.network import CoFiI2P
    from cofii2p_b200.options import Options_KITTI
    from cofii2p_b200.train import TrainStep, training_losses
    from cofii2p_b200.weights import seeded_state_dict
    from oracle import restate
    ops.set_engine(sys.argv[1] if len(sys.argv) > 1 else "fp32")
    opt = Options_KITTI()
    sd = seeded_state_dict(CoFiI2P(opt), 0)
    model = CoFiI2P(opt)
    model.load_state_dict(sd)
    model.cuda()
    frame = get_frame(0, 4096)
    ts = TrainStep(model, opt)
    loss, parts = ts.backward(stack_frames([frame_to(frame, "cuda")]))
    sdr = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k and "kernel_points" not in k
               else v.clone()) for k, v in sd.items()}
    out = restate.forward(sdr, frame["pc_data_dict"], frame["img"], frame["fine_center_kpt_coors"], frame["fine_xy"],
                          frame["fine_pc_inline_index"], "train", bn_training=True)
    sup = {k: frame[k] for k in ("pc_kpt_idx", "pc_outline_idx", "coarse_img_kpt_idx", "K_4", "P", "fine_xy",
                                 "fine_center_kpt_coors")}
    ref_loss, _, _ = training_losses(out, sup, opt, frame["pc_data_dict"]["points"][-1])
    ref_loss.backward()
    print("loss", float(loss), float(ref_loss))
    rows = []
    for k, p in model.named_parameters():
        if p.grad is None or sdr[k].grad is None:
            continue
        g, r = p.grad.cpu().double(), sdr[k].grad.double()
        rows.append(((g - r).norm().item() / max(r.norm().item(), 1e-30), k, r.norm().item(), g.norm().item(),
                     (g - r).abs().max().item(), r.abs().max().item()))
    rows.sort(reverse=True)
    for e, k, rn, gn, mx, rmx in rows[:40]:
        print(f"{e:10.3e} {k:60s} |ref|={rn:.3e} |got|={gn:.3e} maxdiff={mx:.3e} maxref={rmx:.3e}")
    print("order of appearance:")
    names = [k for k, _ in model.named_parameters()]
    bad = [r for r in rows if r[0] > 1e-3]
    for r in sorted(bad, key=lambda r: names.index(r[1])):
        print(f"{r[0]:10.3e} {r[1]}")


if __name__ == "__main__":
    main()
