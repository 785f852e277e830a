#!/usr/bin/env python
"""configs[2] shape: which hand-written calls does one SMOW_Net training step make at a given per-GPU batch, and what does each
cost (eager CUDA events around the C-ABI calls; in-step, operands partly L2-resident)?  Also lists the 25 heaviest CUDA
kernels of the step from the profiler, so that what is NOT hand-written is visible beside it."""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchvision

from smow_net_b200 import ops
from smow_net_b200.models import SMOW_Net
from smow_net_b200.runtime import step as S

B = int(os.environ.get("B", "32"))
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = SMOW_Net(copy.deepcopy(torchvision.models.resnet18(weights=None))).to(dev).train()
a, b = torch.randn(B, 3, 256, 256, device=dev), torch.randn(B, 3, 256, 256, device=dev)
y = (torch.rand(B, 256, 256, device=dev) > 0.95).float()
for _ in range(3):
    model.zero_grad(set_to_none=True)
    S.fwd_bwd(model, a, b, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
with ops.kernel_timer() as kt:
    S.fwd_bwd(model, a, b, y)
e1.record()
torch.cuda.synchronize()
print("step (eager, timed calls): %.2f ms at batch %d" % (e0.elapsed_time(e1), B))
tot = 0.0
for k, v in sorted(kt.summary(by_shape=True).items(), key=lambda kv: -kv[1]["ms"]):
    tot += v["ms"]
    print("  %-40s x%-3d %8.3f ms  %7.0f GB/s" % (k, v["calls"], v["ms"], v["gbps"]))
print("hand-written total: %.2f ms" % tot)
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    S.fwd_bwd(model, a, b, y)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
