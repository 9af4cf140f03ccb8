"""Image stream: ResNet-34 with affine-free InstanceNorm + ImageUpSample decoder
(reference model/imagenet.py:26-73,119-217,360-444) on NHWC tensors and the B200 kernels.

Layout: activations are NHWC ([B,H,W,C] = row-major [pixels, C]) internally, which is the K-major operand
layout of the implicit-GEMM convolution and the token layout of the transformer; public `forward`s accept and
return the reference's NCHW tensors.  `layer3`, `layer4`, `avgpool`, `fc` exist for state_dict compatibility
only: the reference executes them but never uses their outputs (reference model/network.py:87-89)."""
import torch
import torch.nn as nn

from .. import autograd as ad
from .. import ops


def conv3x3(in_planes, out_planes, stride=1, groups=1, dilation=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=dilation, groups=groups, bias=False,
                     dilation=dilation)


def conv1x1(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=1, stride=stride, bias=False)


def packed_conv_weight(conv: nn.Conv2d, cin_pad: int = None):
    """[Cout,Cin,KH,KW] -> [Cout, KH*KW*Cin] (ci fastest), cached per parameter version."""
    w = conv.weight
    v = (w._version, w.data_ptr(), cin_pad, ops.weights_epoch())
    hit = getattr(conv, "_cofi_pack", None)   # cached on the module itself (no global dict keyed by id())
    if hit is not None and hit[0] == v:
        return hit[1]
    with torch.no_grad():
        wp = w.detach().permute(0, 2, 3, 1)
        if cin_pad is not None and cin_pad > wp.shape[3]:
            wp = torch.nn.functional.pad(wp, (0, cin_pad - wp.shape[3]))
        wp = wp.reshape(w.shape[0], -1).contiguous()
    conv._cofi_pack = (v, wp)
    return wp


def run_conv(conv: nn.Conv2d, x_nhwc, scale=None, shift=None, residual=None, act=ops.ACT_NONE, cin_pad=None):
    kh, kw = conv.kernel_size
    if ad.active(conv):  # training: differentiable conv (no folded-BN epilogue on this path)
        assert scale is None and residual is None and act == ops.ACT_NONE
        return ad.conv2d(x_nhwc, conv.weight, conv.stride[0], conv.padding[0])
    return ops.conv2d_nhwc(x_nhwc, packed_conv_weight(conv, cin_pad), kh, kw, conv.stride[0], conv.padding[0],
                           scale=scale, shift=shift, residual=residual, act=act)


def instance_norm_nhwc(x, act=ops.ACT_NONE, residual=None, eps=1e-5):
    """affine-free InstanceNorm2d = per-(image, channel) statistics over H*W rows."""
    B, H, W, C = x.shape
    res = None if residual is None else residual.reshape(B * H * W, C)
    if torch.is_grad_enabled() and x.requires_grad:
        return ad.norm_rows(x.reshape(B * H * W, C), B, C, None, None, eps, res, act).view(B, H, W, C)
    return ops.norm_rows(x.reshape(B * H * W, C), B, C, None, None, eps, residual=res, act=act).view(B, H, W, C)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1, base_width=64, dilation=1,
                 norm_layer=None):
        super().__init__()
        if norm_layer is None:
            norm_layer = nn.BatchNorm2d
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = norm_layer(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = norm_layer(planes)
        self.downsample = downsample
        self.stride = stride

    def forward_nhwc(self, x):
        out = instance_norm_nhwc(run_conv(self.conv1, x), act=ops.ACT_RELU)
        out = run_conv(self.conv2, out)
        identity = x
        if self.downsample is not None:
            identity = instance_norm_nhwc(run_conv(self.downsample[0], x))
        return instance_norm_nhwc(out, act=ops.ACT_RELU, residual=identity)


class ResNet(nn.Module):
    def __init__(self, in_channels, block, layers, num_classes=1000, zero_init_residual=False, groups=1,
                 width_per_group=64, replace_stride_with_dilation=None, norm_layer=nn.InstanceNorm2d):
        super().__init__()
        self._norm_layer = norm_layer
        self.inplanes, self.dilation, self.groups, self.base_width = 64, 1, groups, width_per_group
        self.conv1 = nn.Conv2d(in_channels, self.inplanes, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = norm_layer(self.inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512 * block.expansion, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(conv1x1(self.inplanes, planes * block.expansion, stride),
                                       self._norm_layer(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, downsample, self.groups, self.base_width, 1, self._norm_layer)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes, groups=self.groups, base_width=self.base_width,
                                dilation=self.dilation, norm_layer=self._norm_layer))
        return nn.Sequential(*layers)

    def forward_nhwc(self, img_nchw):
        """-> (s2 [B,80,256,64], s4 [B,40,128,64], s8 [B,20,64,128]) NHWC; the dead layer3/layer4 are skipped."""
        x = ops.nchw_to_nhwc(img_nchw, cpad=4)
        x = instance_norm_nhwc(run_conv(self.conv1, x, cin_pad=4), act=ops.ACT_RELU)
        s2 = x
        x = ad.maxpool2d(x) if (torch.is_grad_enabled() and x.requires_grad) else ops.maxpool2d_3x3s2_nhwc(x)
        for blk in self.layer1:
            x = blk.forward_nhwc(x)
        s4 = x
        for blk in self.layer2:
            x = blk.forward_nhwc(x)
        return s2, s4, x

    def forward(self, x):
        raise RuntimeError("use forward_nhwc (the hot path consumes s2/s4/s8 only)")


def resnet34(in_channels=3, pretrained=False, progress=True, **kwargs):
    return ResNet(in_channels, BasicBlock, [3, 4, 6, 3], **kwargs)


class ImageEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.backbone = resnet34(in_channels=3, pretrained=False, progress=True)

    def forward_nhwc(self, img):
        return self.backbone.forward_nhwc(img)

    def forward(self, x):
        """NCHW [s2, s4, s8] (the live prefix of the reference's 6-element list)."""
        return [ops.nhwc_to_nchw(t) for t in self.forward_nhwc(x)]


def _bn_fold(bn: nn.BatchNorm2d):
    """eval-mode BatchNorm as per-channel (scale, shift), cached per parameter/buffer version."""
    v = (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version,
         bn.weight.data_ptr(), ops.weights_epoch())
    hit = getattr(bn, "_cofi_fold", None)
    if hit is not None and hit[0] == v:
        return hit[1], hit[2]
    with torch.no_grad():
        scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).contiguous()
        shift = (bn.bias - bn.running_mean * scale).contiguous()
    bn._cofi_fold = (v, scale, shift)
    return scale, shift


# forward_batch stacks B independent frames: train-mode BatchNorm then keeps per-frame statistics (the reference
# feeds one frame per call, train.py:226), instead of pooling the stack like a B-image batch would.
PER_FRAME_BN = [False]


def _bn_train(bn: nn.BatchNorm2d, y_nhwc, act, residual=None):
    """train-mode BatchNorm: batch statistics over the H*W rows of each call's batch (+ running-stat update like
    nn.BatchNorm2d); with PER_FRAME_BN every image of the stack is its own batch, updated in order."""
    B, H, W, C = y_nhwc.shape
    res = None if residual is None else residual.reshape(B * H * W, C)
    flat = y_nhwc.reshape(B * H * W, C)
    fr = B if PER_FRAME_BN[0] else 1
    if ad.active(bn):
        out, mean, var = ad.norm_rows(flat, fr, C, bn.weight, bn.bias, bn.eps, res, act, return_stats=True)
    else:
        out, mean, var = ops.norm_rows(flat, fr, C, bn.weight, bn.bias, bn.eps, residual=res, act=act, want_stats=True)
    with torch.no_grad():
        n = B * H * W // fr
        m = bn.momentum
        mean, var = mean.reshape(fr, C), var.reshape(fr, C) * (n / max(n - 1, 1))
        for f in range(fr):
            bn.running_mean.mul_(1 - m).add_(mean[f], alpha=m)
            bn.running_var.mul_(1 - m).add_(var[f], alpha=m)
        bn.num_batches_tracked += fr
    return out.view(B, H, W, C)


class ResidualConv(nn.Module):
    def __init__(self, inplanes, planes, stride=1, kernel_1=False):
        super().__init__()
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = nn.BatchNorm2d(planes)
        if kernel_1:
            self.conv_skip = nn.Sequential(nn.Conv2d(inplanes, planes, kernel_size=1, bias=False), nn.BatchNorm2d(planes))
        else:
            self.conv_skip = nn.Sequential(nn.Conv2d(inplanes, planes, kernel_size=3, stride=stride, padding=1, bias=False),
                                           nn.BatchNorm2d(planes))
        self.stride = stride

    def forward_nhwc(self, x):
        if self.training:
            idt = _bn_train(self.conv_skip[1], run_conv(self.conv_skip[0], x), ops.ACT_NONE)
            out = _bn_train(self.bn1, run_conv(self.conv1, x), ops.ACT_RELU)
            return _bn_train(self.bn2, run_conv(self.conv2, out), ops.ACT_RELU, residual=idt)
        # eval: BatchNorm folded into the conv epilogue (scale/shift), add + ReLU fused as well
        s, t = _bn_fold(self.conv_skip[1])
        idt = run_conv(self.conv_skip[0], x, scale=s, shift=t)
        s, t = _bn_fold(self.bn1)
        out = run_conv(self.conv1, x, scale=s, shift=t, act=ops.ACT_RELU)
        s, t = _bn_fold(self.bn2)
        return run_conv(self.conv2, out, scale=s, shift=t, residual=idt, act=ops.ACT_RELU)


class ImageUpSample(nn.Module):
    def __init__(self, in_channel, output_channel):
        super().__init__()
        self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False)
        self.conv = nn.Sequential(ResidualConv(in_channel, output_channel), ResidualConv(output_channel, output_channel))

    def forward_nhwc(self, x1, x2):
        x = ad.upsample2x_cat(x1, x2) if ad.active(self) else ops.upsample2x_cat_nhwc(x1, x2)
        x = self.conv[0].forward_nhwc(x)
        return self.conv[1].forward_nhwc(x)

    def forward(self, x1, x2):
        return ops.nhwc_to_nchw(self.forward_nhwc(ops.nchw_to_nhwc(x1), ops.nchw_to_nhwc(x2)))
