#!/usr/bin/env python
"""ncu target: flow head forward + backward once at LW's shape (after one warm call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smow_net_b200 import _lib, probe
dev = torch.device("cuda:0"); gen = torch.Generator(device=dev).manual_seed(0)
C = int(os.environ.get("FH_C", "16"))
for name in ("flow_head_fwd", "flow_head_bwd"):
    fn, nbytes, _, keep = probe.build(name, {"B": 16, "C": C, "H": 128, "W": 128, "h": 16, "w": 16}, dev, gen)
    fn(); torch.cuda.synchronize()
    torch.cuda.profiler.start(); fn(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
