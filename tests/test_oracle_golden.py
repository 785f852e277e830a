"""CPU: the oracle (C restatement + torch restatement) against the golden vectors that the real
reference produced (oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

import helpers
from helpers import golden_cases
from oracle import c_oracle, torch_ref


def _scale(a):
    return max(1.0, float(np.abs(a).max()))


@pytest.mark.parametrize("name", golden_cases("warp")[1])
def test_c_oracle_warp_matches_reference(name):
    z, _ = golden_cases("warp")
    g = lambda k: z["warp/%s/%s" % (name, k)]
    out = c_oracle.warp_stack_fwd(g("x"), g("flow"))
    gx, gflow = c_oracle.warp_stack_bwd(g("gout"), g("x"), g("flow"))
    assert np.abs(out - g("out")).max() <= 2e-6
    assert np.abs(gx - g("gx")).max() <= 4e-6 * _scale(g("gx"))
    assert np.abs(gflow - g("gflow")).max() <= 4e-6 * _scale(g("gflow"))
    # pass-through slots are exact copies of the two frames (reference models/SMOW_Net.py:634-636)
    assert np.array_equal(out[:, :, 0], g("x")[:, :, 0]) and np.array_equal(out[:, :, 3], g("x")[:, :, 1])


@pytest.mark.parametrize("name", golden_cases("warp")[1])
def test_torch_restatement_warp_matches_reference(name):
    z, _ = golden_cases("warp")
    g = lambda k: torch.from_numpy(z["warp/%s/%s" % (name, k)])
    out, gx, gflow = torch_ref.warp_with_grads(g("x"), g("flow"), g("gout"))
    # same ATen calls as the reference; only the host's SIMD width can move the last bit
    assert (out - g("out")).abs().max() <= 1e-6
    assert (gx - g("gx")).abs().max() <= 2e-6 * _scale(g("gx").numpy())
    assert (gflow - g("gflow")).abs().max() <= 2e-6 * _scale(g("gflow").numpy())


@pytest.mark.parametrize("name", golden_cases("tlerp")[1])
def test_oracles_tlerp_match_reference(name):
    z, _ = golden_cases("tlerp")
    has_dec = ("tlerp/%s/dec" % name) in z.files
    g = lambda k: z["tlerp/%s/%s" % (name, k)]
    dec = g("dec") if has_dec else None
    cd = dec.shape[1] if has_dec else 0
    cat = c_oracle.tlerp_cat_fwd(dec, g("skip"))
    gskip = c_oracle.tlerp_cat_bwd(g("gcat"), cd)
    assert np.abs(cat - g("cat")).max() <= 1e-6
    assert np.abs(gskip - g("gskip")).max() <= 2e-6
    # frames 0 and 3 are bit-exact copies of T1 / T2
    assert np.array_equal(cat[:, cd:, 0], g("skip")[:, :, 0]) and np.array_equal(cat[:, cd:, 3], g("skip")[:, :, 1])
    t = lambda a: None if a is None else torch.from_numpy(a)
    cat_t, gdec_t, gskip_t = torch_ref.tlerp_cat_with_grads(t(dec), t(g("skip")), t(g("gcat")))
    assert (cat_t - t(g("cat"))).abs().max() <= 1e-6 and (gskip_t - t(g("gskip"))).abs().max() <= 1e-6
    if has_dec:
        assert torch.equal(gdec_t, t(g("gdec")))


def test_zero_flow_is_not_identity():
    """Parity trap 1: with the reference's fp32 linspace grid a zero flow does NOT
    reproduce the input; an 'exact identity' implementation would be wrong."""
    z, _ = golden_cases("warp")
    x, out = z["warp/zero_flow/x"], z["warp/zero_flow/out"]
    dev = np.abs(out[:, :, 1] - x[:, :, 0]).max()
    assert 1e-7 < dev < 1e-3
    mine = c_oracle.warp_stack_fwd(x, z["warp/zero_flow/flow"])
    assert np.abs(mine - out).max() <= 2e-6


def test_border_gradient_gates():
    """clamp mask is inclusive at +-1, the border-clip mask kills the gradient at
    i <= 0 and i >= size-1."""
    z, _ = golden_cases("warp")
    gflow = z["warp/all_clamped/gflow"]
    assert np.all(gflow == 0)
    mine = c_oracle.warp_stack_bwd(z["warp/all_clamped/gout"], z["warp/all_clamped/x"], z["warp/all_clamped/flow"])[1]
    assert np.all(mine == 0)


def test_c_oracle_tokenizer_and_frame_mix_match_the_torch_restatements():
    """Rows N2 / N4: the plain-C restatements (oracle/smow_oracle.c) agree with the ATen-call restatements
    (oracle/torch_ref.py) that tests/test_models_cpu.py pins against the real reference's modules."""
    import torch
    from oracle import c_oracle, torch_ref
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 8, 4, 7, 9, generator=g)
    wa, ba = torch.randn(8, 8, 1, 1, generator=g) * 0.5, torch.randn(8, generator=g)
    tok = c_oracle.tokenizer_fwd(x.numpy(), wa.numpy(), ba.numpy())
    ref = torch_ref.ref_semantic_tokens(x.double(), wa.double(), ba.double()).numpy()
    assert np.abs(tok - ref).max() <= 1e-6
    ws, wo, bias = torch.randn(8, 6, generator=g), torch.randn(4, 8, 6, generator=g), torch.randn(4, 6, generator=g)
    for bb in (None, bias):
        got = c_oracle.frame_mix_fwd(x.numpy(), ws.numpy(), wo.numpy(), None if bb is None else bb.numpy())
        want = torch_ref.ref_cyclic_frame_mix(x.double(), ws.double(), wo.double(), None if bb is None else bb.double()).numpy()
        assert np.abs(got - want).max() <= 1e-5


def test_flow_head_restatement_reproduces_the_reference_golden():
    """Row N1: oracle/torch_ref.py::ref_flow_head (interpolate + cat + conv3d, reference models/SMOW_Net.py:606-608) against
    tests/golden/flow_head.npz, whose `flow` was recorded from the REAL reference's OFW.forward (oracle/make_golden.py
    ::gen_flow_head) — forward bit for bit, autograd gradients to fp32 rounding."""
    import torch
    from oracle import torch_ref
    z = helpers.load_golden("flow_head.npz")
    names = sorted({k.split("/")[1] for k in z.files})
    assert names
    for n in names:
        t = {k: torch.from_numpy(z["flow_head/%s/%s" % (n, k)]) for k in ("x", "coarse", "weight", "gflow", "flow", "gx", "gcoarse", "gweight")}
        x, c, w = (t[k].clone().requires_grad_(True) for k in ("x", "coarse", "weight"))
        flow = torch_ref.ref_flow_head(x, c, w)
        assert torch.equal(flow.detach(), t["flow"]), n
        flow.backward(t["gflow"])
        for got, want in ((x.grad, t["gx"]), (c.grad, t["gcoarse"]), (w.grad, t["gweight"])):
            assert float((got - want).abs().max()) <= 1e-5 * max(1.0, float(want.abs().max())), n
