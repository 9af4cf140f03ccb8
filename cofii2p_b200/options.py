"""Configuration bag read by `CoFiI2P(opt)`: the attributes the reference model consumes
(reference data/options.py:17-22,53; model/network.py:18-27), KITTI values."""
import torch


class Options_KITTI:
    def __init__(self):
        self.img_H = 160
        self.img_W = 512
        self.img_fine_resolution_scale = 32
        self.num_pc = 20480
        self.num_kpt = 64
        self.norm = "gn"
        self.group_norm = 32
        # training (reference data/options.py:39-58): correspondence radius at 1/8 resolution, circle-loss margins, Adam
        self.dist_thres = 1.0
        self.pos_margin = 0.2
        self.neg_margin = 1.8
        self.lr = 1e-3
        self.min_lr = 1e-5
        self.lr_decay_step = 0.25
        self.lr_decay_scale = 0.5
        self.device = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
