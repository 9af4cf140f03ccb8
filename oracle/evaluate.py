"""TEST INFRASTRUCTURE ONLY -- CPU restatement of reference evaluation/eval_all.py:99-105 (fine matching and pixel
coordinate assembly), line for line in tensor algebra, applied to test-mode outputs (the reference's own frozen outputs in
tests/golden/*.npz, or oracle/restate.py's).  The PnP-RANSAC call of :107 is OpenCV in the reference and here alike."""
import torch


def correspondences(fine_img_feature_patch, fine_pc_inline_feature, fine_center_xy, coarse_pc_points):
    f = fine_pc_inline_feature.unsqueeze(-1)                                                   # :99 (after the :98 rearrange: [n,C,16])
    patch = fine_img_feature_patch.reshape(fine_img_feature_patch.shape[0], fine_img_feature_patch.shape[1], -1)
    dist = torch.cosine_similarity(patch.unsqueeze(-1), f.unsqueeze(-2))                       # :100
    dist = torch.squeeze(dist)                                                                 # :101
    predict_index = torch.argmax(dist, dim=1)                                                  # :102
    fine_xy = fine_center_xy - 2                                                               # :103
    fine_xy[0] = fine_xy[0] + predict_index // 4                                               # :104
    fine_xy[1] = fine_xy[1] + predict_index % 4                                                # :105
    return fine_xy.T.numpy(), coarse_pc_points.numpy(), predict_index
