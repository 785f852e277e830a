"""GPU: the drop-in modules end to end on the CUDA kernels, against the real reference's golden
outputs (generated on the CPU by oracle/model_fixture.py) and against the oracle-routed module on the
same device.  TF32 is disabled so cuDNN/cuBLAS stay comparable with the CPU reference."""
import numpy as np
import pytest
import torch

import helpers
from oracle.model_fixture import run_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def strict_fp32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


class _OnDevice(torch.nn.Module):
    """run_model() works on CPU tensors; this wrapper moves I/O so the fixture code is shared."""

    def __init__(self, m):
        super().__init__()
        self.m = m

    def forward(self, a, b):
        return self.m(a.to(DEV), b.to(DEV)).cpu()

    def named_parameters(self, *a, **k):
        return ((n, _CpuGrad(p)) for n, p in self.m.named_parameters())

    def named_buffers(self, *a, **k):
        return ((n, b.detach().cpu()) for n, b in self.m.named_buffers())


class _CpuGrad:
    def __init__(self, p):
        self.grad = None if p.grad is None else p.grad.detach().cpu()


@pytest.mark.parametrize("kind", ["lw", "s"])
def test_module_on_cuda_matches_reference_golden(kind):
    from smow_net_b200 import _lib
    model = helpers.seeded_model(kind, device=DEV)
    x1, x2 = helpers.seeded_pair(2)
    before = _lib.launch_count()
    got = run_model(_OnDevice(model), x1, x2, helpers.seeded_labels(2))
    assert _lib.launch_count() - before >= 14   # the hand-written kernels really ran
    want = helpers.load_golden("model_%s.npz" % kind)
    for k in want.files:
        a, b = want[k], got[k]
        if a.dtype == np.uint8:
            diff = int(np.unpackbits(a ^ b).sum())
            assert diff <= 0.001 * a.size * 8, (k, diff)          # change maps: <= 0.1 % of pixels differ
        elif k.startswith("grad/") or k == "loss" or k.startswith("bn/"):
            err = np.abs(a.astype(np.float64) - b).max() / max(1e-12, np.abs(a).max())
            assert err < 2e-3, (k, err)                            # different conv algorithms CPU vs cuDNN
        else:
            assert np.abs(a - b).max() <= 1e-4, (k, np.abs(a - b).max())


@pytest.mark.parametrize("kind", ["lw", "s"])
def test_module_matches_oracle_routed_module_on_same_device(kind, monkeypatch):
    """Same weights, same device, same cuDNN: only the hot-path operators differ."""
    model = helpers.seeded_model(kind, device=DEV).eval()
    x1, x2 = (t.to(DEV) for t in helpers.seeded_pair(2, seed=3))
    with torch.no_grad():
        mine = model(x1, x2)
    helpers.use_oracle_ops(monkeypatch)
    with torch.no_grad():
        ref = model(x1, x2)
    assert float((mine - ref).abs().max()) <= 1e-5
    assert float(((mine > 0.5) != (ref > 0.5)).float().mean()) <= 1e-3


def test_tiled_inference_matches_per_crop_calls():
    """Config 4: a 512x768 scene tile is cut into 256x256 crops (the only size the reference network accepts),
    run as one batch and re-assembled; every crop equals a direct call on that crop."""
    from smow_net_b200.runtime import synthetic
    model = helpers.seeded_model("lw", device=DEV).eval()
    g = torch.Generator().manual_seed(17)
    ta, tb = torch.randn(1, 3, 512, 768, generator=g).to(DEV), torch.randn(1, 3, 512, 768, generator=g).to(DEV)
    with torch.no_grad():
        prob = model(synthetic.tiles_to_crops(ta), synthetic.tiles_to_crops(tb))
        full = synthetic.crops_to_tiles(prob, 1, 512, 768)
        direct = model(ta[:, :, 256:512, 512:768].contiguous(), tb[:, :, 256:512, 512:768].contiguous())
    assert full.shape == (1, 1, 512, 768)
    assert float((full[:, :, 256:512, 512:768] - direct).abs().max()) <= 1e-5


@pytest.mark.parametrize("with_optimizer", [False, True])
def test_graphed_step_replays_the_eager_step(with_optimizer):
    """runtime/graph.py: one CUDA graph of forward + loss + backward (+ clip + AdamW) gives the eager step's loss and
    gradients on fresh inputs, and the hand-written kernels are inside the graph."""
    from smow_net_b200.runtime import graph as G, step as S, synthetic
    torch.manual_seed(5)
    model = helpers.seeded_model("lw", device=DEV).train()
    S.freeze_unused(model)
    a, b, y = synthetic.make_batch(2, device=DEV, seed=11)
    a2, b2, y2 = synthetic.make_batch(2, device=DEV, seed=12)
    twin = helpers.seeded_model("lw", device=DEV).train()
    S.freeze_unused(twin)
    opt = S.make_optimizer(model, capturable=True) if with_optimizer else None
    opt2 = S.make_optimizer(twin) if with_optimizer else None
    gs = G.GraphedStep(model, a, b, y, optimizer=opt, warmup=2)
    assert gs.hot_path_launches >= 14
    if with_optimizer:
        # the twin takes the same number of eager steps on the same batches (capture itself executes nothing)
        for _ in range(2):
            S.train_step(twin, opt2, None, a, b, y)
        loss_g = float(gs(a2, b2, y2).detach())
        loss_e = float(S.train_step(twin, opt2, None, a2, b2, y2).detach())
        assert abs(loss_g - loss_e) <= 2e-3 * max(1.0, abs(loss_e))
        pa, pb = dict(model.named_parameters()), dict(twin.named_parameters())
        worst = max(float((pa[k] - pb[k]).abs().max()) for k in pa)
        assert worst <= 5e-3
    else:
        twin.load_state_dict(model.state_dict())       # BatchNorm statistics moved during warm-up + capture
        loss_g = float(gs(a2, b2, y2).detach())
        loss_e = float(S.fwd_bwd(twin, a2, b2, y2).detach())
        assert abs(loss_g - loss_e) <= 1e-5 * max(1.0, abs(loss_e))
        ga = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
        gb = {k: p.grad for k, p in twin.named_parameters() if p.grad is not None}
        assert ga.keys() == gb.keys()
        for k in ga:
            scale = max(1.0, float(gb[k].abs().max()))
            assert float((ga[k] - gb[k]).abs().max()) <= 2e-4 * scale, k


@pytest.mark.parametrize("case", [("conv", 24, 24, 9, 7), ("deconv_bias", 20, 12, 9, 7),          # cuBLAS formulation
                                  ("conv", 16, 16, 9, 7), ("conv", 16, 16, 128, 128), ("deconv_bias", 28, 28, 40, 37),
                                  ("conv", 32, 32, 64, 64), ("deconv_bias", 64, 64, 32, 33), ("conv", 28, 28, 5, 3)])
def test_cyclic_frame_mix_matches_the_composed_reference(case):
    """Row N4: the hand-written frame-mix kernels (C = 16 / 28 / 32 / 64) and the GEMM formulation (other channel counts)
    equal the reference's composition of slices, 1x1x1 convolutions, adds and a concat (models/SMOW_Net.py:121-139) —
    outputs, input gradient and all parameter gradients."""
    from smow_net_b200 import _lib
    from smow_net_b200.models import blocks
    kind, cin, cout, H, W = case
    torch.manual_seed(3)
    mk = (lambda: torch.nn.Conv3d(cin, cout, 1, bias=False)) if kind == "conv" else \
         (lambda: torch.nn.ConvTranspose3d(cin, cout, 1, bias=True))
    mods = [mk().to(DEV) for _ in range(5)]
    B = 2 if H * W > 4096 else 3
    x = torch.randn(B, cin, 4, H, W, device=DEV).contiguous(memory_format=torch.channels_last_3d)
    g = torch.randn(B, cout, 4, H, W, device=DEV)
    res, launches = [], []
    for fn in (blocks.cyclic_frame_mix, blocks._cyclic_frame_mix_composed):
        for m in mods:
            m.zero_grad(set_to_none=True)
        xi = x.clone(memory_format=torch.preserve_format).requires_grad_(True)
        before = _lib.launch_count()
        y = fn(xi, mods[4], mods[:4])
        y.backward(g)
        launches.append(_lib.launch_count() - before)
        res.append([y.detach(), xi.grad] + [p.grad.clone() for m in mods for p in m.parameters()])
    fused = cin == cout and cin in (16, 28, 32, 64)
    assert launches == [4 if fused else 0, 0]          # apply fwd + apply bwd + wgrad (2 kernels)
    assert res[0][0].is_contiguous(memory_format=torch.channels_last_3d)
    for a, b in zip(*res):
        assert a.shape == b.shape
        assert float((a - b).abs().max()) <= 2e-5 * max(1.0, float(b.abs().max()))
    if fused:       # fixed summation order: bit-reproducible
        for m in mods:
            m.zero_grad(set_to_none=True)
        xi = x.clone(memory_format=torch.preserve_format).requires_grad_(True)
        y = blocks.cyclic_frame_mix(xi, mods[4], mods[:4])
        y.backward(g)
        again = [y.detach(), xi.grad] + [p.grad.clone() for m in mods for p in m.parameters()]
        assert all(torch.equal(a, b) for a, b in zip(res[0], again))
