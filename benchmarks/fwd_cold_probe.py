#!/usr/bin/env python
"""NDHWC warp forward at the configs[4] shapes, fp32 and bf16 storage: HBM-cold graph replays (>= 1 GiB rotation)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smow_net_b200 import _lib, probe
dev = torch.device("cuda:0")
for dt, name in ((_lib.F32, "f32"), (_lib.BF16, "bf16")):
    res = []
    for C, H in ((64, 64), (64, 128), (64, 256), (128, 128), (256, 128), (256, 256), (16, 128), (32, 128)):
        s = 4 if dt == _lib.F32 else 2
        B = max(1, min(256, int((1 << 30) / (6 * C * H * H * s))))
        m = {"B": B, "C": C, "H": H, "W": H, "dtype": dt, "layout": _lib.NDHWC, "pair": 0}
        t = probe.time_call("warp_stack_fwd", m, dev, footprint=1 << 30, max_sets=4, sigma=0.3)
        res.append("C%d/%d B%d %.3f ms (%.2f)" % (C, H, B, t["cold_ms"], t["bytes"] / t["cold_ms"] / 1e6 / 6547.8))
    print(name + ": " + " | ".join(res), flush=True)
