from .kpconv import KPConv
from .modules import ConvBlock, ResidualBlock, UnaryBlock, LastUnaryBlock, GroupNorm, MaxPool
from .functional import nearest_upsample, maxpool
from .kp_backbone import KPConvFPN
