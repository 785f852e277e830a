#!/usr/bin/env python
"""ncu target: the fused BatchNorm + LeakyReLU + lerp backward (reduce, finalize, apply) once, cold.
BN_ROWS=0 selects the split dec / skip apply kernel; BN_CASES="Cd:Cs,Cd:Cs" the channel splits (default LW's largest level)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smow_net_b200 import _lib, probe
dev = torch.device("cuda:0"); gen = torch.Generator(device=dev).manual_seed(0)
F32, ND = _lib.F32, _lib.NDHWC
_lib.set_option("bn_bwd_rows", int(os.environ.get("BN_ROWS", "1")))
for case in os.environ.get("BN_CASES", "28:16").split(","):
    cd, cs = (int(v) for v in case.split(":"))
    meta = {"B": 16, "Cd": cd, "Cs": cs, "hw": 16384, "dtype": F32, "layout": ND, "pair": 0, "act": 2}
    fn, nbytes, _, keep = probe.build("tlerp_cat_bwd", meta, dev, gen)
    fn(); torch.cuda.synchronize()
    torch.cuda.profiler.start(); fn(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
    del fn, keep; torch.cuda.empty_cache()
