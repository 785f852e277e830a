"""Decoder / projection blocks shared by SMOW_Net and SMOW_Net_LW (out of the hot path, cuDNN).

Parameter names follow the reference so its checkpoints load strictly
(models/SMOW_Net.py:104-159,411-423 and models/SMOW_Net_LW.py:102-178); the code is organised
around one idea the reference spells out four times: a spatial (1,k,k) convolution whose four
output frames are then mixed *cyclically* by 1x1x1 convolutions,

    out[j] = shared(frame[j]) + own[j+1](frame[(j+1) % 4]),

``shared`` being initialised to the identity and every ``own`` to zero.
"""
import torch
import torch.nn as nn


def cyclic_frame_mix(frames5d, shared, own):
    """frames5d: (B,C,4,H,W); shared: module; own: list of 4 modules (own[k] acts on frame k)."""
    parts = [frames5d[:, :, k:k + 1] for k in range(4)]
    kept = [shared(p) for p in parts]
    lent = [m(p) for m, p in zip(own, parts)]
    return torch.cat([kept[j] + lent[(j + 1) % 4] for j in range(4)], dim=2)


def _identity_1x1(conv):
    with torch.no_grad():
        torch.nn.init.eye_(conv.weight[:, :, 0, 0, 0])


def _zero(conv):
    torch.nn.init.constant_(conv.weight, 0.0)


class TemporalDeconvMix(nn.Module):
    """x2 spatial transposed conv + cyclic frame mixing + BN + LeakyReLU(0.2).

    ``wide=True`` is SMOW_Net's block (5x5 kernel, biased time convs, models/SMOW_Net.py:104-139);
    ``wide=False`` is SMOW_Net_LW's (3x3 kernel, bias-free, models/SMOW_Net_LW.py:102-137)."""

    def __init__(self, in_dim, out_dim, wide=True):
        super().__init__()
        k = 5 if wide else 3
        self.conv3d_spatial = nn.ConvTranspose3d(in_dim, out_dim, kernel_size=(1, k, k), stride=(1, 2, 2),
                                                 padding=(0, k // 2, k // 2), output_padding=(0, 1, 1))
        t_in = in_dim if wide else out_dim
        for i in range(1, 6):
            setattr(self, "conv3d_time_%d" % i, nn.ConvTranspose3d(t_in, out_dim, kernel_size=1, bias=wide))
        for i in range(1, 5):
            _zero(getattr(self, "conv3d_time_%d" % i))
        _identity_1x1(self.conv3d_time_5)
        self.batch = nn.BatchNorm3d(out_dim)
        self.leaky = nn.LeakyReLU(0.2, inplace=True)

    def forward(self, x):
        own = [getattr(self, "conv3d_time_%d" % i) for i in range(1, 5)]
        y = cyclic_frame_mix(self.conv3d_spatial(x), self.conv3d_time_5, own)
        return self.leaky(self.batch(y))


class DoubleConv3d(nn.Module):
    """SMOW_Net's conv_block_2_3d: two 3x3x3 convs, BN after each, LeakyReLU between
    (models/SMOW_Net.py:142-156).  The Sequential keeps the reference's attribute name."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.conv_block_2_3d = nn.Sequential(
            nn.Conv3d(in_dim, out_dim, kernel_size=3, padding=1), nn.BatchNorm3d(out_dim),
            nn.LeakyReLU(0.2, inplace=True),
            nn.Conv3d(out_dim, out_dim, kernel_size=3, padding=1), nn.BatchNorm3d(out_dim))

    def forward(self, x):
        return self.conv_block_2_3d(x)


class SpatialConvMix(nn.Module):
    """SMOW_Net_LW's conv_block_2_3d: (1,3,3) conv + cyclic frame mixing + BN + LeakyReLU
    (models/SMOW_Net_LW.py:140-175)."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.conv3d_s = nn.Conv3d(in_dim, out_dim, kernel_size=(1, 3, 3), padding=(0, 1, 1))
        for i in range(1, 6):
            setattr(self, "conv3d_t%d" % i, nn.Conv3d(out_dim, out_dim, kernel_size=1, bias=False))
        for i in range(1, 5):
            _zero(getattr(self, "conv3d_t%d" % i))
        _identity_1x1(self.conv3d_t5)
        self.b = nn.BatchNorm3d(out_dim)
        self.l = nn.LeakyReLU(0.2, inplace=True)

    def forward(self, x):
        own = [getattr(self, "conv3d_t%d" % i) for i in range(1, 5)]
        y = cyclic_frame_mix(self.conv3d_s(x), self.conv3d_t5, own)
        return self.l(self.b(y))


class PointwiseConvBN(nn.Module):
    """BasicConv3d: 1x1x1 conv + BN + ReLU under ``conv_bn`` (models/SMOW_Net.py:411-423)."""

    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv_bn = nn.Sequential(nn.Conv3d(in_ch, out_ch, kernel_size=1), nn.BatchNorm3d(out_ch), nn.ReLU())

    def forward(self, x):
        return self.conv_bn(x)


def convs_channels_last_3d(module):
    """Store every 5-D (3-D convolution) weight of `module` in channels_last_3d order, in place.  cuDNN then runs
    the decoder / encoder convolutions and BatchNorms as NDHWC kernels: 52.7 -> 42.7 ms (SMOW_Net) and
    38.1 -> 33.6 ms (SMOW_Net_LW) per batch-16 step on B200 (benchmarks/e2e_probe.py).  Shapes, names and values
    of the parameters are untouched, so state_dicts stay interchangeable with the reference."""
    for p in module.parameters():
        if p.dim() == 5:
            p.data = p.data.contiguous(memory_format=torch.channels_last_3d)
    return module


def spatial_max_pool():
    """max_pooling_3d(): halves H and W, keeps the 4 frames (models/SMOW_Net.py:158-159)."""
    return nn.MaxPool3d(kernel_size=(1, 2, 2), stride=(1, 2, 2), padding=0)
