"""The reference's own operator sequence (ATen kernels), restated for TIMING comparisons in benchmarks/.
Kept separate from oracle/ on purpose: oracle/ is reserved for tests, smoke() and bench.py's CPU legs."""
import torch
import torch.nn.functional as F


def aten_flow_warp(input, flow):
    """ATen sequence of OFW.flow_warp (reference models/SMOW_Net.py:612-638)."""
    B, C, T, H, W = input.shape
    ys = torch.linspace(-1.0, 1.0, H).view(-1, 1).expand(H, W)
    xs = torch.linspace(-1.0, 1.0, W).view(1, -1).expand(H, W)
    grid = torch.stack((xs, ys), 2).unsqueeze(0).expand(B, H, W, 2).type_as(input).to(input.device)
    norm = torch.tensor([[[[W, H]]]]).type_as(input).to(input.device)
    frames = []
    for t in range(T):
        field = flow[:, :, t].permute(0, 2, 3, 1) / norm
        frames.append(F.grid_sample(input[:, :, t], (grid + field).clamp(-1, 1), mode="bilinear",
                                    padding_mode="border", align_corners=True).unsqueeze(2))
    return torch.cat([input[:, :, 0:1]] + frames + [input[:, :, 1:2]], dim=2)


def aten_tlerp_cat(dec, skip):
    """ATen sequence of the temporal up-sample + skip concat (reference models/SMOW_Net.py:64-73,78-94)."""
    b, c, t, h, w = skip.shape
    up = F.interpolate(skip, size=(4, h, w), mode="trilinear", align_corners=True)
    return up if dec is None else torch.cat([dec, up], dim=1)
