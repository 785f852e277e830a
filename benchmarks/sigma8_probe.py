#!/usr/bin/env python
"""warp backward at flow sigma 8 (every tap far): tile gather + far pass vs the vector-atomic scatter."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smow_net_b200 import _lib, probe
dev = torch.device("cuda:0")
for dt in (_lib.F32, _lib.BF16):
    for C, H, B in ((64, 64, 127), (128, 128, 16), (256, 128, 8), (16, 128, 64)):
        m = {"B": B, "C": C, "H": H, "W": H, "dtype": dt, "layout": _lib.NDHWC, "pair": 0}
        res = []
        for sigma in (0.3, 8.0):
            for var in ((-1, 0) if dt == _lib.F32 else (-1,)):
                _lib.set_option("warp_bwd_variant", var)
                t = probe.time_call("warp_stack_bwd", m, dev, footprint=1 << 30, max_sets=4, sigma=sigma)
                res.append("s%.1f v%d %.3f ms (%.2f)" % (sigma, var, t["cold_ms"], t["bytes"] / t["cold_ms"] / 1e6 / 6547.8))
        _lib.set_option("warp_bwd_variant", -1)
        print("dtype %d C%d H%d B%d: " % (dt, C, H, B) + " | ".join(res), flush=True)
