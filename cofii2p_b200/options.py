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
        self.device = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
