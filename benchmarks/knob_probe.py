#!/usr/bin/env python
"""Quick A/B of tuning knobs on one HBM-cold shape (NDHWC backward)."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smow_net_b200 import _lib, ops
dev = "cuda:0"; B, C, H = 64, 32, 128
g = torch.Generator(device=dev).manual_seed(0)
cl = torch.channels_last_3d
x = torch.randn(B, C, 2, H, H, device=dev, generator=g).contiguous(memory_format=cl).requires_grad_(True)
flow = (torch.randn(B, 2, 2, H, H, device=dev, generator=g) * 0.3).requires_grad_(True)
gout = torch.randn(B, C, 4, H, H, device=dev, generator=g).contiguous(memory_format=cl)
def t(fn, n=10, inner=4):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(inner): fn()
        b.record(); b.synchronize(); ts.append(a.elapsed_time(b) / inner)
    return statistics.median(ts)
out = ops.flow_warp(x, flow, (H, H))
def bwd():
    x.grad = None; flow.grad = None; out.backward(gout, retain_graph=True)
for mb in (0, 64, 128, 256, 512):
    _lib.set_option("bwd_chunk_mb", mb)
    print("chunk_mb", mb, "bwd %.3f ms" % t(bwd))
_lib.set_option("bwd_chunk_mb", 0); _lib.set_option("warp_bwd_variant", 3)
print("deterministic gather bwd %.3f ms" % t(bwd))
