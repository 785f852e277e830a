"""ResNet3D — a torchvision ResNet-18 re-used as a bi-temporal (T=2) 3-D encoder (cuDNN; not hot path).

Mirrors the structure and parameter names of the reference (models/SMOW_Net.py:426-585): every
2-D conv becomes a (1,k,k) ``conv3d_spatial`` that re-uses the 2-D weights, followed by three 1x1x1
"temporal exchange" convs (``conv3d_time_2`` = identity on the own frame, ``_1``/``_3`` = zero-initialised
cross-frame terms); BatchNorm2d modules are kept as they are (so ImageNet statistics load) and only
taught to accept 5-D input; down-sample shortcuts become plain (1,1,1)-in-time 3-D convs.
"""
import torch
import torch.nn as nn


def _lift_weight(conv2d):
    return nn.Parameter(conv2d.weight.data.unsqueeze(2))


class BiTemporalConv(nn.Module):
    """Decompose_conv with time_dim=3 (models/SMOW_Net.py:426-473): spatial conv, then
    out[T1] = time_2(T1) + time_3(T2),  out[T2] = time_1(T1) + time_2(T2)."""

    def __init__(self, conv2d):
        super().__init__()
        ch = conv2d.out_channels
        self.conv3d_spatial = nn.Conv3d(conv2d.in_channels, ch, kernel_size=(1,) + tuple(conv2d.kernel_size),
                                        stride=(1, conv2d.stride[0], conv2d.stride[0]),
                                        padding=(0,) + tuple(conv2d.padding),
                                        dilation=(1,) + tuple(conv2d.dilation))
        self.conv3d_spatial.weight = _lift_weight(conv2d)
        self.conv3d_spatial.bias = conv2d.bias
        for i in (1, 2, 3):
            setattr(self, "conv3d_time_%d" % i, nn.Conv3d(ch, ch, kernel_size=1, bias=False))
        nn.init.constant_(self.conv3d_time_1.weight, 0.0)
        nn.init.constant_(self.conv3d_time_3.weight, 0.0)
        with torch.no_grad():
            nn.init.eye_(self.conv3d_time_2.weight[:, :, 0, 0, 0])

    def forward(self, x):
        # row N4, encoder flavour: T = 2 exchange  out[T1] = time_2(T1) + time_3(T2),  out[T2] = time_2(T2) + time_1(T1)
        # (one tcgen05 pass on the GPU instead of two slices, four 1x1x1 convolutions, two adds and a concat)
        from .blocks import cyclic_frame_mix
        return cyclic_frame_mix(self.conv3d_spatial(x), self.conv3d_time_2, [self.conv3d_time_3, self.conv3d_time_1],
                                shift=1, own_off=0)


def accept_5d(bn2d):
    """Decompose_norm (models/SMOW_Net.py:475-478): keep the 2-D BatchNorm object, drop its rank check."""
    bn2d._check_input_dim = nn.BatchNorm3d(bn2d.num_features)._check_input_dim
    return bn2d


def lift_pool(pool2d):
    """Decompose_pool with time_dim=1 for the stem max-pool (models/SMOW_Net.py:480-497)."""
    return nn.MaxPool3d((1, pool2d.kernel_size, pool2d.kernel_size), stride=(1, pool2d.stride, pool2d.stride),
                        padding=(0, pool2d.padding, pool2d.padding), dilation=(1, pool2d.dilation, pool2d.dilation),
                        ceil_mode=pool2d.ceil_mode)


def lift_shortcut(downsample2d):
    """Decompose_downsample (models/SMOW_Net.py:581-585): 1x1 stride-s conv -> (1,1,1) 3-D conv, + BN."""
    c = downsample2d[0]
    conv = nn.Conv3d(c.in_channels, c.out_channels, kernel_size=(1,) + tuple(c.kernel_size),
                     stride=(1, c.stride[0], c.stride[0]), padding=(0,) + tuple(c.padding),
                     dilation=(1,) + tuple(c.dilation))
    conv.weight = _lift_weight(c)
    conv.bias = c.bias
    return nn.Sequential(conv, accept_5d(downsample2d[1]))


class BiTemporalBlock(nn.Module):
    """Bottleneck3d wrapping a torchvision BasicBlock (models/SMOW_Net.py:542-579).
    NB: like the reference, a ReLU follows bn2 *before* the residual add."""

    def __init__(self, block2d):
        super().__init__()
        self.conv1 = BiTemporalConv(block2d.conv1)
        self.bn1 = accept_5d(block2d.bn1)
        self.conv2 = BiTemporalConv(block2d.conv2)
        self.bn2 = accept_5d(block2d.bn2)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = lift_shortcut(block2d.downsample) if block2d.downsample is not None else None
        self.stride = block2d.stride

    def forward(self, x):
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.relu(self.bn2(self.conv2(y)))
        shortcut = x if self.downsample is None else self.downsample(x)
        return self.relu(y + shortcut)


class ResNet3D(nn.Module):
    def __init__(self, resnet2d):
        super().__init__()
        self.conv1 = BiTemporalConv(resnet2d.conv1)
        self.bn1 = accept_5d(resnet2d.bn1)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = lift_pool(resnet2d.maxpool)
        for name in ("layer1", "layer2", "layer3", "layer4"):
            setattr(self, name, nn.Sequential(*[BiTemporalBlock(b) for b in getattr(resnet2d, name)]))
