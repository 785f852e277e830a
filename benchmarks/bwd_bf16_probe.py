#!/usr/bin/env python
"""warp backward, bf16 storage on NDHWC (tile gather): HBM-cold graph replays at the sweep's shapes x tile height."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smow_net_b200 import _lib, probe
dev = torch.device("cuda:0")
ROWS = tuple(int(a) for a in sys.argv[1].split(",")) if len(sys.argv) > 1 else (0, 4, 8)
SHAPES = ((64, 128, 64), (128, 128, 32), (256, 128, 16), (64, 256, 16), (64, 64, 128), (32, 128, 64), (16, 128, 64))
for C, H, B in SHAPES:
    m = {"B": B, "C": C, "H": H, "W": H, "dtype": _lib.BF16, "layout": _lib.NDHWC, "pair": 0}
    res = []
    for rows in ROWS:
        _lib.set_option("ndhwc_bwd_rows", rows)
        try:
            t = probe.time_call("warp_stack_bwd", m, dev, footprint=1 << 30, max_sets=4, sigma=0.3)
            res.append("R%d %.3f ms (%.2f)" % (rows, t["cold_ms"], t["bytes"] / t["cold_ms"] / 1e6 / 6547.8))
        except RuntimeError:
            res.append("R%d does not fit" % rows)
    print("bf16 C%d H%d B%d: " % (C, H, B) + " | ".join(res), flush=True)
_lib.set_option("ndhwc_bwd_rows", 0)
