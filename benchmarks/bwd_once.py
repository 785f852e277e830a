#!/usr/bin/env python
"""ncu target: a few NDHWC warp backward launches on one HBM-cold shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smow_net_b200 import _lib, ops
B, C, H = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (64, 32, 128)
sigma = float(sys.argv[4]) if len(sys.argv) > 4 else 0.3
dev, cl = "cuda:0", torch.channels_last_3d
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B, C, 2, H, H, device=dev, generator=g).contiguous(memory_format=cl).requires_grad_(True)
flow = (torch.randn(B, 2, 2, H, H, device=dev, generator=g) * sigma).requires_grad_(True)
gout = torch.randn(B, C, 4, H, H, device=dev, generator=g).contiguous(memory_format=cl)
out = ops.flow_warp(x, flow, (H, H))
for _ in range(3):
    x.grad = None; flow.grad = None
    out.backward(gout, retain_graph=True)
torch.cuda.synchronize()
print("ok")
