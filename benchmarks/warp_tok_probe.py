#!/usr/bin/env python
"""Rows A1 + N2: fused warp -> tokens pass (ops.warp_tokens; the stack never reaches HBM) against the two launches it
replaces (warp_stack_fwd + tokenizer_fwd; backward: tokenizer_bwd vs the re-staging backward).  HBM-cold and warm
graph-replayed timings of the C-ABI calls (smow_net_b200.probe)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from smow_net_b200 import _lib, probe

dev = torch.device("cuda:0")
F32, ND = _lib.F32, _lib.NDHWC
for B, C in ((16, 16), (16, 32), (128, 32)):
    H = W = 128
    t = {}
    for name, meta in (("warp_stack_fwd", {"B": B, "C": C, "H": H, "W": W, "dtype": F32, "layout": ND, "pair": 0}),
                       ("tokenizer_fwd", {"B": B, "C": C, "hw": H * W}),
                       ("tokenizer_bwd", {"B": B, "C": C, "hw": H * W}),
                       ("warp_tokens_fwd", {"B": B, "C": C, "H": H, "W": W}),
                       ("warp_tokens_bwd", {"B": B, "C": C, "H": H, "W": W})):
        r = probe.time_call(name, meta, dev)
        t[name] = r
        print("B%d C%d %-16s %7.1f us cold (%5.0f GB/s algorithmic) %7.1f us warm" % (
            B, C, name, r["cold_ms"] * 1e3, r["bytes"] / r["cold_ms"] / 1e6, r["warm_ms"] * 1e3), flush=True)
        torch.cuda.empty_cache()
    print("B%d C%d forward: fused %.1f us vs warp + tokenizer %.1f us; backward pooling half: fused %.1f us vs %.1f us" % (
        B, C, t["warp_tokens_fwd"]["cold_ms"] * 1e3, (t["warp_stack_fwd"]["cold_ms"] + t["tokenizer_fwd"]["cold_ms"]) * 1e3,
        t["warp_tokens_bwd"]["cold_ms"] * 1e3, t["tokenizer_bwd"]["cold_ms"] * 1e3), flush=True)
