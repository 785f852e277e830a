"""CPU: runtime/step.py and runtime/metrics.py against the REFERENCE's own training utilities
(utils/loss_f.py:8-18, utils/func.py:4-8, utils/metric_tool.py:93-155, train.py:162-185), loaded unmodified through
oracle/ref_runtime.py (oracle/_ref/reference.zip, built by oracle/build_ref.py)."""
import copy

import numpy as np
import pytest
import torch

from oracle import build_ref, ref_runtime
from smow_net_b200.runtime import metrics, step as S


@pytest.fixture(scope="module")
def ref():
    build_ref.build()
    m = ref_runtime.modules()
    if m is None:
        pytest.skip("oracle/_ref/reference.zip absent and no reference tree to build it from")
    return m


def test_bce_dice_loss_equals_the_reference_loss(ref):
    g = torch.Generator().manual_seed(0)
    for shape in ((2, 16, 16), (1, 5, 7), (3, 64, 64)):
        pred = torch.rand(shape, generator=g).clamp(1e-4, 1 - 1e-4)
        true = (torch.rand(shape, generator=g) > 0.9).float()
        p1, p2 = pred.clone().requires_grad_(True), pred.clone().requires_grad_(True)
        mine, want = S.bce_dice_loss(p1, true), ref["loss_f"].BCEDICE_loss(p2, true)
        mine.backward()
        want.backward()
        assert torch.equal(mine, want)
        assert torch.equal(p1.grad, p2.grad)


def test_clip_gradient_equals_the_reference_clamp(ref):
    g = torch.Generator().manual_seed(1)
    params = [torch.nn.Parameter(torch.randn(n, generator=g)) for n in (3, 17, 256)] + [torch.nn.Parameter(torch.zeros(2))]
    for p in params[:3]:
        p.grad = torch.randn(p.shape, generator=g) * 2
    theirs = copy.deepcopy(params)
    for p, q in zip(params, theirs):
        q.grad = None if p.grad is None else p.grad.clone()
    S.clip_gradient_(params, 0.5)
    ref["func"].clip_gradient(torch.optim.SGD(theirs, 0.1), 0.5)
    for p, q in zip(params, theirs):
        assert (p.grad is None and q.grad is None) or torch.equal(p.grad, q.grad)
    assert float(params[0].grad.abs().max()) <= 0.5


def _tiny():
    torch.manual_seed(5)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.BatchNorm2d(4), torch.nn.ReLU(),
                               torch.nn.Conv2d(4, 1, 1), torch.nn.Sigmoid())


class _Pair(torch.nn.Module):
    def __init__(self, net):
        super().__init__()
        self.net = net

    def forward(self, a, b):
        return self.net(a - b)


def test_train_step_follows_the_reference_step_sequence(ref):
    """train.py:162-179: zero_grad, forward, squeeze, BCEDICE_loss, backward, clip_gradient, AdamW.step, scheduler.step."""
    mine, theirs = _Pair(_tiny()), _Pair(_tiny())
    theirs.load_state_dict(mine.state_dict())
    opt_m = S.make_optimizer(mine)
    sch_m = S.make_scheduler(opt_m, 6)
    opt_t = torch.optim.AdamW(theirs.parameters(), 1e-4, weight_decay=1e-4)                       # train.py:135
    sch_t = torch.optim.lr_scheduler.CosineAnnealingLR(opt_t, T_max=6, eta_min=1e-6)            # utils/lr_scheduler.py:65-69
    g = torch.Generator().manual_seed(3)
    for it in range(4):
        a, b = torch.randn(2, 3, 8, 8, generator=g), torch.randn(2, 3, 8, 8, generator=g)
        y = (torch.rand(2, 8, 8, generator=g) > 0.8).float()
        l_m = S.train_step(mine, opt_m, sch_m, a, b, y)
        opt_t.zero_grad()
        pred = theirs(a, b).squeeze(1)
        l_t = ref["loss_f"].BCEDICE_loss(pred, y)
        l_t.backward()
        ref["func"].clip_gradient(opt_t, 0.5)
        opt_t.step()
        sch_t.step()
        assert torch.equal(l_m, l_t), it
    for (n, p), (_, q) in zip(mine.state_dict().items(), theirs.state_dict().items()):
        assert torch.equal(p, q), n
    assert opt_m.param_groups[0]["lr"] == opt_t.param_groups[0]["lr"]


def test_confusion_meter_equals_the_reference_meter(ref):
    g = torch.Generator().manual_seed(9)
    meter = metrics.ConfusionMeter("cpu")
    theirs = ref["metric_tool"].ConfuseMatrixMeter(n_class=2)
    for _ in range(3):
        pred = torch.rand(4, 32, 32, generator=g)
        gts = (torch.rand(4, 32, 32, generator=g) > 0.7).float()
        meter.update(pred, gts)
        theirs.update_cm(pr=(pred > 0.5).numpy().astype(int), gt=gts.numpy().astype(int))      # train.py:182-185
    assert np.array_equal(meter.cm.numpy(), theirs.sum.astype(np.int64))
    want, got = theirs.get_scores(), meter.scores()
    for k in ("acc", "iou", "F1", "precision", "recall"):
        assert abs(got[k] - float(want[k])) <= 1e-12, k
    meter.reset()
    assert int(meter.cm.sum()) == 0
