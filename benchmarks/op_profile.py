#!/usr/bin/env python
"""torch.profiler table of one eager step (which ATen ops own the GPU time)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from smow_net_b200.runtime import launch, step as S, synthetic
kind = sys.argv[1] if len(sys.argv) > 1 else "lw"
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
model = launch.build_model(kind, dev).train()
a, b, y = synthetic.make_batch(16, device=dev, seed=1)
for _ in range(5): S.fwd_bwd(model, a, b, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): S.fwd_bwd(model, a, b, y)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=45, max_name_column_width=60))
