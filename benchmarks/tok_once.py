#!/usr/bin/env python
"""ncu target: semantic tokenizer forward + backward on one shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smow_net_b200 import ops
B, C, H = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (64, 16, 128)
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B, C, 4, H, H, device=dev, generator=g).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
w = (torch.randn(8, C, 1, 1, device=dev, generator=g) / C ** 0.5).requires_grad_(True)
b = torch.randn(8, device=dev, generator=g).requires_grad_(True)
gt = torch.randn(B, 4, 8, C, device=dev, generator=g)
for _ in range(2):
    x.grad = None
    ops.semantic_tokens(x, w, b).backward(gt)
torch.cuda.synchronize()
print("ok")
