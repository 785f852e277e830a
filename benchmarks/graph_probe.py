#!/usr/bin/env python
"""Eager vs CUDA-graph step time (SMOW_Net_LW / SMOW_Net fwd+bwd, batch 16)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smow_net_b200.runtime import graph as G, launch, step as S, synthetic
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
for kind, B in (("lw", 16), ("s", 16)):
    model = launch.build_model(kind, dev).train()
    a, b, y = synthetic.make_batch(B, device=dev, seed=1)
    for _ in range(5): S.fwd_bwd(model, a, b, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(20): S.fwd_bwd(model, a, b, y)
    e1.record(); torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / 20
    gs = G.GraphedStep(model, a, b, y, warmup=3)
    for _ in range(3): gs()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20): gs()
    e1.record(); torch.cuda.synchronize()
    graph = e0.elapsed_time(e1) / 20
    print("%s B=%d eager %.2f ms  graph %.2f ms  (%.0f -> %.0f pairs/s), hot-path launches/replay %d" %
          (kind, B, eager, graph, B / eager * 1e3, B / graph * 1e3, gs.hot_path_launches), flush=True)
