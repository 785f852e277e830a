"""OFW — optical-flow warp module (rows A1/A2 of SURVEY §8).

Same constructor, parameters and ``state_dict`` keys as the reference's ``OFW``
(models/SMOW_Net.py:587-638, duplicated at models/SMOW_Net_LW.py:429-480):
``down.{0,3,6}`` depthwise stride-(1,2,2) convs with BatchNorm ``down.{1,4,7}`` and the
bias-free ``flow_make`` head.  The depthwise ``down`` stack stays on cuDNN; the rest of the flow head (x8 up-sample,
concat, ``flow_make``) is one hand-written pass (``ops.flow_head``, row N1); ``flow_warp`` — base grid,
clamp, two bilinear border-mode samplings and the [T1, warp(T1), warp(T2), T2] stack — is
one launch of the sm_100a kernel behind ``smow_net_b200.ops.flow_warp``.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


def _down_stage(ch):
    return [nn.Conv3d(ch, ch, kernel_size=3, stride=(1, 2, 2), padding=1, groups=ch),
            nn.BatchNorm3d(ch), nn.ReLU(inplace=True)]


class OFW(nn.Module):
    def __init__(self, inplane):
        super().__init__()
        self.down = nn.Sequential(*(_down_stage(inplane) + _down_stage(inplane) + _down_stage(inplane)))
        self.flow_make = nn.Conv3d(2 * inplane, 2, kernel_size=3, padding=1, bias=False)

    def predict_flow(self, x):
        """(B,C,2,H,W) -> flow (B,2,2,H,W); reference models/SMOW_Net.py:606-608.

        The reference hard-codes the upsample target (2,128,128), i.e. 256x256 network inputs;
        here it is the size of ``x`` itself, which is the same thing at 256x256."""
        coarse = self.down(x)
        if ops.flow_head_supported(x, coarse):
            # row N1: up-sample + concat + flow_make as one bandwidth-bound pass (no up-sampled tensor, no concat)
            return ops.flow_head(x, coarse, self.flow_make.weight)
        coarse = F.interpolate(coarse, size=tuple(x.shape[2:]), mode="trilinear", align_corners=True)
        return self.flow_make(torch.cat([x, coarse], dim=1))

    def flow_warp(self, input, flow, size):
        return ops.flow_warp(input, flow, size)

    def forward(self, x):
        return self.flow_warp(x, self.predict_flow(x), x.shape[3:])
