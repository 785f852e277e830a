#!/usr/bin/env python
"""One forward + backward of the flow-head kernels at the models' shapes (for an ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smow_net_b200 import _lib, probe
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
for B, C in ((16, 16), (16, 32)):
    m = {"B": B, "C": C, "H": 128, "W": 128, "h": 16, "w": 16}
    for name in ("flow_head_fwd", "flow_head_bwd"):
        fn = probe.build(name, m, dev, gen)[0]
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
print("done")
