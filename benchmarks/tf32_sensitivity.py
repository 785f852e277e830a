#!/usr/bin/env python
"""How far do the REFERENCE-path gradients move between cuDNN TF32 on and off (no hand-written kernel involved)?
Both arms: the drop-in module with every operator of the path routed to the oracle (ATen)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import helpers
from smow_net_b200.runtime import step as S

class MP:
    def setattr(self, obj, name, val): setattr(obj, name, val)

helpers.use_oracle_ops(MP())
DEV = "cuda:0"
for kind in ("lw", "s"):
    x1, x2 = (t.to(DEV) for t in helpers.seeded_pair(2, seed=21))
    y = helpers.seeded_labels(2, seed=22).to(DEV)
    grads = []
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        m = helpers.seeded_model(kind, device=DEV).train()
        loss, pred = S.forward_loss(m, x1, x2, y)
        loss.backward()
        grads.append(({k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}, pred.detach(), float(loss)))
    (ga, pa, la), (gb, pb, lb) = grads
    rel = sorted(((float((ga[k] - gb[k]).norm()) / float(gb[k].norm()), k) for k in gb if float(gb[k].norm()) > 1e-6), reverse=True)
    print(kind, "pred max diff %.2e, loss %.6f vs %.6f; grad rel-L2 worst:" % (float((pa - pb).abs().max()), la, lb), [(round(r, 3), k) for r, k in rel[:4]])
