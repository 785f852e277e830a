#!/usr/bin/env python
"""GPU-only time of the NDHWC warp backward at the models' in-step shapes (batch 16): the C-ABI call is captured in a
CUDA graph and replayed, so Python / launch latency is out of the picture.  A/B of the tile height knob."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smow_net_b200 import _lib, ops
dev, cl = torch.device("cuda:0"), torch.channels_last_3d
lib = _lib.load()


def replay_ms(fn, n=50):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): g.replay()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n / 10


for B, C, H in ((16, 16, 128), (16, 32, 128), (128, 32, 128)):
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(B, C, 2, H, H, device=dev, generator=gen).contiguous(memory_format=cl)
    flow = torch.randn(B, 2, 2, H, H, device=dev, generator=gen) * 0.3
    gout = torch.randn(B, C, 4, H, H, device=dev, generator=gen).contiguous(memory_format=cl)
    out = torch.empty_like(gout); gx = torch.empty_like(x); gflow = torch.empty_like(flow)
    xs, ys = ops.base_grid(H, dev), ops.base_grid(H, dev)
    ws = torch.zeros(64, dtype=torch.uint8, device=dev)

    def bwd():
        _lib.check(lib.smow_warp_stack_bwd(gout.data_ptr(), x.data_ptr(), flow.data_ptr(), xs.data_ptr(), ys.data_ptr(),
                                           gx.data_ptr(), gflow.data_ptr(), B, C, H, H, _lib.F32, _lib.NDHWC,
                                           ws.data_ptr(), 64, torch.cuda.current_stream().cuda_stream), "bwd")

    def fwd():
        _lib.check(lib.smow_warp_stack_fwd(x.data_ptr(), flow.data_ptr(), xs.data_ptr(), ys.data_ptr(), out.data_ptr(),
                                           B, C, H, H, _lib.F32, _lib.NDHWC, torch.cuda.current_stream().cuda_stream), "fwd")
    nb, nf = ops.warp_bwd_bytes(B, C, H, H, 4), ops.warp_fwd_bytes(B, C, H, H, 4)
    res = ["fwd %.1f us %.0f GB/s" % (replay_ms(fwd) * 1e3, nf / replay_ms(fwd) / 1e6)]
    for rows in (0, 4, 6, 8, 12):
        _lib.set_option("ndhwc_bwd_rows", rows)
        ms = replay_ms(bwd)
        res.append("bwd[R=%d] %.1f us %.0f GB/s" % (rows, ms * 1e3, nb / ms / 1e6))
    _lib.set_option("ndhwc_bwd_rows", 0)
    _lib.set_option("warp_bwd_variant", 0)
    ms = replay_ms(bwd); res.append("scatter %.1f us" % (ms * 1e3))
    _lib.set_option("warp_bwd_variant", -1)
    print("B%d C%d H%d: " % (B, C, H) + " | ".join(res), flush=True)
