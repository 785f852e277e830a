#!/usr/bin/env python
"""warp backward (fp32 NDHWC tile gather), HBM-cold graph replays at the sweep's shapes: tile height x tap prefetch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smow_net_b200 import _lib, probe
dev = torch.device("cuda:0")
ROWS = tuple(int(a) for a in sys.argv[1].split(",")) if len(sys.argv) > 1 else (0, 2, 4, 8)
for C, H, B in ((16, 128, 16), (16, 128, 64), (32, 128, 64), (64, 64, 127), (64, 256, 8), (128, 128, 16), (256, 128, 8)):
    m = {"B": B, "C": C, "H": H, "W": H, "dtype": _lib.F32, "layout": _lib.NDHWC, "pair": 0}
    for pf in (0, -1):
        _lib.set_option("ndhwc_bwd_pf", pf)
        res = []
        for rows in ROWS:
            _lib.set_option("ndhwc_bwd_rows", rows)
            t = probe.time_call("warp_stack_bwd", m, dev, footprint=1 << 30, max_sets=4, sigma=0.3)
            res.append("R%d %.3f ms (%.2f)" % (rows, t["cold_ms"], t["bytes"] / t["cold_ms"] / 1e6 / 6547.8))
        print("C%d H%d B%d pf %2d: " % (C, H, B, pf) + " | ".join(res), flush=True)
_lib.set_option("ndhwc_bwd_rows", 0)
_lib.set_option("ndhwc_bwd_pf", -1)
