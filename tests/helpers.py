"""Shared test utilities: golden loading, seeded models, oracle-backed operator substitution."""
import copy
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = os.environ.get("SMOW_REFERENCE", "/root/reference")


def load_golden(name="hotpath_ops.npz"):
    return np.load(os.path.join(GOLDEN, name))


def golden_cases(prefix):
    z = load_golden()
    names = sorted({k.split("/")[1] for k in z.files if k.startswith(prefix + "/")})
    return z, names


def seeded_model(kind, seed=1234, device="cpu"):
    """A SMOW_Net ('s') or SMOW_Net_LW ('lw') whose every parameter and BatchNorm statistic is a
    deterministic function of `seed` (CPU generator), with the zero-initialised temporal-exchange
    convs perturbed so that every branch of the network is live."""
    import torchvision
    from smow_net_b200.models import SMOW_Net, SMOW_Net_LW
    torch.manual_seed(seed)
    if kind == "s":
        model = SMOW_Net(copy.deepcopy(torchvision.models.resnet18(weights=None)))
    else:
        model = SMOW_Net_LW(pretrained=False)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, p in model.named_parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.02)
        for name, b in model.named_buffers():
            if name.endswith("running_mean"):
                b.copy_(torch.randn(b.shape, generator=g) * 0.1)
            elif name.endswith("running_var"):
                b.copy_(1.0 + 0.2 * torch.rand(b.shape, generator=g))
    return model.to(device)


def seeded_pair(batch, seed=99, size=256):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 3, size, size, generator=g), torch.randn(batch, 3, size, size, generator=g)


def seeded_labels(batch, seed=7, size=256):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(batch, size, size, generator=g) > 0.95).float()


def use_oracle_ops(monkeypatch):
    """Route the four hot-path operators to the PyTorch restatement of the reference (oracle/torch_ref.py).
    TESTS ONLY: lets the module plumbing be checked on a CPU-only box; the product has no such path."""
    from oracle import torch_ref
    from smow_net_b200 import ops
    monkeypatch.setattr(ops, "flow_warp", lambda x, flow, size=None: torch_ref.ref_flow_warp(x, flow))
    monkeypatch.setattr(ops, "tlerp", torch_ref.ref_tlerp)
    monkeypatch.setattr(ops, "tlerp_cat", torch_ref.ref_tlerp_cat)
    monkeypatch.setattr(ops, "tlerp_pair_cat",
                        lambda dec, a, b: torch_ref.ref_tlerp_cat(dec, torch_ref.ref_pair_stack(a, b)))
    # the fused activation + concat forms: the reference's own LeakyReLU (models/SMOW_Net.py:137), then interpolate + cat
    leaky = torch.nn.functional.leaky_relu
    monkeypatch.setattr(ops, "act_tlerp_cat", lambda z, skip, slope=0.2: torch_ref.ref_tlerp_cat(leaky(z, slope), skip))
    monkeypatch.setattr(ops, "act_tlerp_pair_cat",
                        lambda z, a, b, slope=0.2: torch_ref.ref_tlerp_cat(leaky(z, slope), torch_ref.ref_pair_stack(a, b)))
    monkeypatch.setattr(ops, "act_cat_supported", lambda z, cs: True)
    # rows N2 / N4: the tokenizer and the cyclic frame mix go to the reference's op sequence too, so a module-level
    # comparison never has the CUDA kernels of those rows in both arms
    from smow_net_b200.models import blocks
    monkeypatch.setattr(ops, "semantic_tokens", torch_ref.ref_semantic_tokens)
    monkeypatch.setattr(ops, "flow_head", torch_ref.ref_flow_head)            # row N1
    # rows A1 + N2 fused: the reference's own sequence, flow_warp then the pooling
    monkeypatch.setattr(ops, "warp_tokens",
                        lambda x, flow, w, b: torch_ref.ref_semantic_tokens(torch_ref.ref_flow_warp(x, flow), w, b))

    def mix(frames5d, shared, own, shift=1, own_off=1):
        T, bias = len(own), None
        if shared.bias is not None:
            bias = torch.stack([shared.bias + own[(f + own_off) % T].bias for f in range(T)])
        return torch_ref.ref_cyclic_frame_mix(frames5d, blocks._mix_matrix(shared),
                                              torch.stack([blocks._mix_matrix(m) for m in own]), bias, shift, own_off)
    monkeypatch.setattr(blocks, "cyclic_frame_mix", mix)
    monkeypatch.setattr(blocks, "fused_bn_enabled", lambda *a, **k: False)     # BatchNorm stays the framework's in the oracle arm


def import_reference():
    """The real reference modules, or None where /root/reference does not exist (the GPU box)."""
    import sys
    if not os.path.isdir(REFERENCE):
        return None
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import models.SMOW_Net as ref_s
    import models.SMOW_Net_LW as ref_lw
    ref_lw.load_state_dict_from_url = lambda *a, **k: {}
    return ref_s, ref_lw


def bce_dice(pred, true):
    """utils/loss_f.py:8-18 of the reference: BCE + (1 - global dice), eps 1e-7."""
    bce = torch.nn.functional.binary_cross_entropy(pred, true)
    inter = (pred * true).sum()
    dice = (2.0 * inter + 1e-7) / (pred.sum() + true.sum() + 1e-7)
    return bce + 1.0 - dice
