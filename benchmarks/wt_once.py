#!/usr/bin/env python
"""ncu target: fused warp -> tokens forward + backward once, cold (B = 64, C = 32, 128 x 128)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smow_net_b200 import _lib, probe
dev = torch.device("cuda:0"); gen = torch.Generator(device=dev).manual_seed(0)
for name in ("warp_tokens_fwd", "warp_tokens_bwd"):
    fn, nbytes, _, keep = probe.build(name, {"B": 64, "C": 32, "H": 128, "W": 128}, dev, gen)
    fn(); torch.cuda.synchronize()
    torch.cuda.profiler.start(); fn(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
    del fn, keep; torch.cuda.empty_cache()
