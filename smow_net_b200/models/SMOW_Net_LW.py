"""SMOW_Net_LW — drop-in for the reference's light-weight model (models/SMOW_Net_LW.py:9-99).

Same ``forward(x1, x2)``, attribute names and ``state_dict`` keys (482 entries).  The reference
constructor takes no argument and downloads ImageNet weights for its MobileNetV2; here
``SMOW_Net_LW()`` does the same, and ``SMOW_Net_LW(pretrained=False)`` skips the download
(benchmarks and tests run without network, with random-init weights).

Hot-path differences: the OFW warp + stack is one launch; the Siamese pyramids are
never stacked for levels 1..4 — ``ops.tlerp_pair_cat`` reads the T1 and T2 feature maps through
two pointers and writes the 4-frame interpolation directly into the decoder's concat buffer
(rows A3, A4, A5).  Level 0 is stacked once because OFW's 3-D convs need the (B,C,2,H,W) tensor.
"""
import torch
import torch.nn as nn

from .. import ops
from .blocks import SpatialConvMix, TemporalDeconvMix, convs_channels_last_3d, spatial_max_pool
from .mobilenet import mobilenet_v2
from .ofw import OFW
from .tokens import Classifier, Transformer_Decoder, Transformer_Encoder

# decoder level k: (C3DTk channels, C3Dk in, C3Dk out); skip of level k is pyramid level 5-k
_DECODER = {1: (320, 640, 160), 2: (160, 256, 64), 3: (64, 96, 32), 4: (32, 56, 28), 5: (28, 44, 16)}


class SMOW_Net_LW(nn.Module):
    def __init__(self, pretrained=True):
        super().__init__()
        # The 2-D Siamese backbone runs channels-last on the GPU: cuDNN's NHWC depthwise / pointwise kernels
        # are ~2x faster on B200 than ATen's NCHW depthwise ones (benchmarks/e2e_probe.py: 44.3 -> 39.0 ms per
        # batch-16 step).  Parameter shapes, names and values are unchanged.
        self.backbone = mobilenet_v2(pretrained=pretrained).to(memory_format=torch.channels_last)
        self.OFW = OFW(16)
        self.MaxPool = spatial_max_pool()
        for k, (cdec, ccat, cout) in _DECODER.items():
            setattr(self, "C3DT%d" % k, TemporalDeconvMix(cdec, cdec, wide=False))
            setattr(self, "C3D%d" % k, SpatialConvMix(ccat, cout))
        self.Transformer_Encoder = Transformer_Encoder(in_chan=16)
        self.Transformer_Decoder = Transformer_Decoder(in_chan=64)
        self.decoder = Classifier(in_chan=64, n_class=1)
        self.sigmoid = nn.Sigmoid()
        convs_channels_last_3d(self)

    def forward(self, x1, x2):
        if x1.is_cuda:
            x1 = x1.contiguous(memory_format=torch.channels_last)
            x2 = x2.contiguous(memory_format=torch.channels_last)
        pyr1 = self.backbone(x1)                                # 5 levels, T1
        pyr2 = self.backbone(x2)                                # same weights, T2

        # (B,16,2,H/2,W/2), reference :38-40 — stacked straight into channels_last_3d on the GPU so that OFW's
        # convolutions and the warp kernels all see NDHWC
        x0 = torch.empty((pyr1[0].shape[0], pyr1[0].shape[1], 2) + tuple(pyr1[0].shape[2:]), dtype=pyr1[0].dtype,
                         device=pyr1[0].device,
                         memory_format=torch.channels_last_3d if x1.is_cuda else torch.contiguous_format)
        x0[:, :, 0] = pyr1[0]
        x0[:, :, 1] = pyr2[0]
        tokens = self.Transformer_Encoder.from_warp(self.OFW, x0)

        dec = self.MaxPool(ops.tlerp_pair_cat(None, pyr1[4], pyr2[4]))   # reference :71-73
        for k in (1, 2, 3, 4):
            cat = getattr(self, "C3DT%d" % k).forward_into_concat(dec, skip_pair=(pyr1[5 - k], pyr2[5 - k]))
            dec = getattr(self, "C3D%d" % k)(cat)
        dec = self.C3D5(self.C3DT5.forward_into_concat(dec, skip=x0))

        return self.sigmoid(self.decoder(self.Transformer_Decoder(dec, tokens)))
