#!/usr/bin/env python
"""Quick A/B of NDHWC backward strategies / tuning knobs on HBM-cold and in-step shapes."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smow_net_b200 import _lib, ops
dev = "cuda:0"
cl = torch.channels_last_3d


def t(fn, n=10, inner=4):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(inner): fn()
        b.record(); b.synchronize(); ts.append(a.elapsed_time(b) / inner)
    return statistics.median(ts)


for (B, C, H, sigma) in ((64, 32, 128, 0.3), (121, 16, 128, 0.3), (16, 16, 128, 0.3), (16, 128, 128, 0.3), (8, 64, 256, 0.3), (64, 32, 128, 1.5), (64, 32, 128, 8.0)):
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(B, C, 2, H, H, device=dev, generator=g).contiguous(memory_format=cl).requires_grad_(True)
    flow = (torch.randn(B, 2, 2, H, H, device=dev, generator=g) * sigma).requires_grad_(True)
    gout = torch.randn(B, C, 4, H, H, device=dev, generator=g).contiguous(memory_format=cl)
    out = ops.flow_warp(x, flow, (H, H))
    nbytes = ops.warp_bwd_bytes(B, C, H, H, 4)

    def bwd():
        x.grad = None; flow.grad = None; out.backward(gout, retain_graph=True)
    res = []
    _lib.set_option("warp_bwd_variant", 0)
    ms = t(bwd); res.append("scatter %.3f ms %.0f GB/s" % (ms, nbytes / ms / 1e6))
    _lib.set_option("warp_bwd_variant", -1)
    for rows in (0, 8, 12, 16):
        _lib.set_option("ndhwc_bwd_rows", rows)
        ms = t(bwd); res.append("tile[R=%d] %.3f ms %.0f GB/s" % (rows, ms, nbytes / ms / 1e6))
    _lib.set_option("ndhwc_bwd_rows", 0)
    print("B%d C%d H%d sigma %.1f: " % (B, C, H, sigma) + " | ".join(res), flush=True)
