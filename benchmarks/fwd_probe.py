#!/usr/bin/env python
"""NDHWC warp forward: shuffled per-pixel coordinates (default) vs per-(pixel, vector) kernel (variant 0)."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smow_net_b200 import _lib, ops
dev, cl = "cuda:0", torch.channels_last_3d


def t(fn, n=10, inner=4):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(inner): fn()
        b.record(); b.synchronize(); ts.append(a.elapsed_time(b) / inner)
    return statistics.median(ts)


for B, C, H in ((121, 16, 128), (63, 32, 128), (32, 64, 128), (8, 256, 128), (16, 16, 128)):
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(B, C, 2, H, H, device=dev, generator=g).contiguous(memory_format=cl)
    flow = torch.randn(B, 2, 2, H, H, device=dev, generator=g) * 0.3
    nb = ops.warp_fwd_bytes(B, C, H, H, 4)
    res = []
    with torch.no_grad():
        for v in (-1, 0):
            _lib.set_option("warp_fwd_variant", v)
            ms = t(lambda: ops.flow_warp(x, flow, (H, H)))
            res.append("variant %d: %.3f ms %.0f GB/s" % (v, ms, nb / ms / 1e6))
    _lib.set_option("warp_fwd_variant", -1)
    print("B%d C%d H%d: " % (B, C, H) + " | ".join(res), flush=True)
