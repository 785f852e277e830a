"""SMOW_Net — drop-in for the reference module of the same name (models/SMOW_Net.py:8-101).

Same constructor (``SMOW_Net(resnet18)`` takes a torchvision ResNet-18), same
``forward(x1, x2) -> (B,1,H,W)`` sigmoid change map, same sub-module attribute names and
``state_dict`` keys (411 entries).  What changes is the hot path:

* ``OFW``: the warp + [T1, warp(T1), warp(T2), T2] stack is one sm_100a kernel launch;
* the five ``F.interpolate(x_k, size=(4,h,w), 'trilinear')`` + ``torch.cat([dec, x_k], 1)`` pairs
  (:64-73, :78-94) are one launch each (``ops.tlerp_cat``): the up-sampled skip is written
  straight into the concat buffer and never exists on its own.

Everything else (ResNet3D encoder, decoder convs, token transformers) is cuDNN/cuBLAS.
"""
import torch
import torch.nn as nn

from .. import ops
from .blocks import DoubleConv3d, PointwiseConvBN, TemporalDeconvMix, convs_channels_last_3d, spatial_max_pool
from .encoder3d import ResNet3D
from .ofw import OFW
from .tokens import Classifier, Transformer_Decoder, Transformer_Encoder

# decoder level k = 1..5: (C3DTk in/out, C3Dk in, C3Dk out); skip of level k is x_{5-k}
_DECODER = {1: (256, 512, 128), 2: (128, 256, 64), 3: (64, 128, 64), 4: (64, 96, 32), 5: (32, 64, 32)}
# projection of encoder stage s -> skip channels
_PROJ = {"Conv3d": (64, 32), "Conv3d1": (64, 32), "Conv3d2": (128, 64), "Conv3d3": (256, 128), "Conv3d4": (512, 256)}


class SMOW_Net(nn.Module):
    def __init__(self, resnet18):
        super().__init__()
        self.resnet = ResNet3D(resnet18)
        self.OFW = OFW(32)
        for name, (cin, cout) in _PROJ.items():
            setattr(self, name, PointwiseConvBN(cin, cout))
        self.MaxPool = spatial_max_pool()
        for k, (cdec, ccat, cout) in _DECODER.items():
            setattr(self, "C3DT%d" % k, TemporalDeconvMix(cdec, cdec, wide=True))
            setattr(self, "C3D%d" % k, DoubleConv3d(ccat, cout))
        self.Transformer_Encoder = Transformer_Encoder(in_chan=32)
        self.Transformer_Decoder = Transformer_Decoder(in_chan=128)
        self.decoder = Classifier(in_chan=128, n_class=1)
        self.sigmoid = nn.Sigmoid()
        convs_channels_last_3d(self)

    def forward(self, x1, x2):
        x = torch.stack((x1, x2), dim=2)                       # (B,3,2,H,W), reference :40-42
        stem = self.resnet.relu(self.resnet.bn1(self.resnet.conv1(x)))
        skips = [self.Conv3d(stem)]                            # x0 (B,32,2,H/2,W/2)
        tokens = self.Transformer_Encoder.from_warp(self.OFW, skips[0])  # warp -> semantic tokens

        feat = self.resnet.maxpool(stem)
        for s in (1, 2, 3, 4):
            feat = getattr(self.resnet, "layer%d" % s)(feat)
            skips.append(getattr(self, "Conv3d%d" % s)(feat))  # x1..x4

        dec = self.MaxPool(ops.tlerp(skips[4]))                # reference :73-75
        for k in (1, 2, 3, 4, 5):
            cat = getattr(self, "C3DT%d" % k).forward_into_concat(dec, skip=skips[5 - k])
            dec = getattr(self, "C3D%d" % k)(cat)

        return self.sigmoid(self.decoder(self.Transformer_Decoder(dec, tokens)))
