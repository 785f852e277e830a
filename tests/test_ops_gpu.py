"""GPU: the CUDA kernels, called through the C ABI (smow_net_b200.ops -> libsmow_b200.so), against
  (1) golden vectors produced by the real reference,
  (2) the oracle's torch restatement of the reference run on the same device (= ATen's own CUDA path),
  (3) size-independent properties at BASELINE.json's full sizes.
Tolerance (north_star): max-abs <= 1e-5 in fp32 for unit-scale tensors (gradients whose magnitude
exceeds 1 are compared relative to their max), <= 2e-2 for bf16 storage."""
import numpy as np
import pytest
import torch

from helpers import golden_cases
from oracle import torch_ref
from smow_net_b200 import _lib, ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FWD_VARIANTS = [0, 1, 2]
BWD_VARIANTS = [0, 1, 2]


def _scale(t):
    return max(1.0, float(t.abs().max()))


@pytest.fixture
def variants():
    saved = {k: _lib.get_option(k) for k in ("warp_fwd_variant", "warp_bwd_variant")}
    yield
    for k, v in saved.items():
        _lib.set_option(k, v)


def run_warp(x, flow, gout):
    x = x.detach().clone().requires_grad_(True)
    flow = flow.detach().clone().requires_grad_(True)
    out = ops.flow_warp(x, flow, x.shape[3:])
    out.backward(gout)
    return out.detach(), x.grad.detach(), flow.grad.detach()


def check_warp(got, want, tol=1e-5):
    for name, a, b in zip(("out", "gx", "gflow"), got, want):
        err = float((a.float() - b.float()).abs().max())
        assert err <= tol * _scale(b.float()), "%s: max-abs error %.3e (scale %.2f)" % (name, err, _scale(b.float()))


@pytest.mark.parametrize("bv", BWD_VARIANTS)
@pytest.mark.parametrize("fv", FWD_VARIANTS)
@pytest.mark.parametrize("name", golden_cases("warp")[1])
def test_warp_against_reference_golden(name, fv, bv, variants):
    _lib.set_option("warp_fwd_variant", fv)
    _lib.set_option("warp_bwd_variant", bv)
    z, _ = golden_cases("warp")
    g = lambda k: torch.from_numpy(z["warp/%s/%s" % (name, k)]).to(DEV)
    got = run_warp(g("x"), g("flow"), g("gout"))
    check_warp(got, (g("out"), g("gx"), g("gflow")))
    assert torch.equal(got[0][:, :, 0], g("x")[:, :, 0]) and torch.equal(got[0][:, :, 3], g("x")[:, :, 1])


# (B, C, H, W, sigma): the two models' OFW shapes, sweep corners, ragged shapes, large displacement
FULL_CASES = [(4, 32, 128, 128, 0.3), (4, 16, 128, 128, 0.3), (2, 64, 64, 64, 8.0), (1, 128, 256, 256, 0.3),
              (2, 256, 64, 64, 0.3), (3, 8, 40, 72, 2.0), (2, 5, 17, 23, 1.0), (1, 4, 128, 128, 40.0),
              (2, 12, 100, 128, 3.0)]


@pytest.mark.parametrize("bv", BWD_VARIANTS)
@pytest.mark.parametrize("fv", FWD_VARIANTS)
@pytest.mark.parametrize("case", FULL_CASES)
def test_warp_against_same_device_reference(case, fv, bv, variants):
    _lib.set_option("warp_fwd_variant", fv)
    _lib.set_option("warp_bwd_variant", bv)
    B, C, H, W, sigma = case
    g = torch.Generator(device=DEV).manual_seed(B * 1000 + C)
    x = torch.randn(B, C, 2, H, W, device=DEV, generator=g)
    flow = torch.randn(B, 2, 2, H, W, device=DEV, generator=g) * sigma
    gout = torch.randn(B, C, 4, H, W, device=DEV, generator=g)
    check_warp(run_warp(x, flow, gout), torch_ref.warp_with_grads(x, flow, gout))


@pytest.mark.parametrize("fv", FWD_VARIANTS)
def test_warp_forward_is_bit_exact_vs_aten_cuda(fv, variants):
    """The coordinate chain is reproduced bit for bit, and the 4-tap accumulation uses ATen's order."""
    _lib.set_option("warp_fwd_variant", fv)
    g = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(2, 32, 2, 128, 128, device=DEV, generator=g)
    flow = torch.randn(2, 2, 2, 128, 128, device=DEV, generator=g) * 1.5
    with torch.no_grad():
        mine, ref = ops.flow_warp(x, flow, (128, 128)), torch_ref.ref_flow_warp(x, flow)
    assert float((mine - ref).abs().max()) <= 1e-6
    assert float((mine != ref).float().mean()) < 1e-3   # essentially every element identical


@pytest.mark.parametrize("bv", BWD_VARIANTS)
@pytest.mark.parametrize("fv", FWD_VARIANTS)
def test_warp_bf16_storage(fv, bv, variants):
    """bf16 features, fp32 flow/coordinates/accumulation; oracle = fp32 reference on bf16-rounded
    features (SURVEY §7 'bf16': the reference's own bf16 path quantises the grid and is unusable)."""
    _lib.set_option("warp_fwd_variant", fv)
    _lib.set_option("warp_bwd_variant", bv)
    g = torch.Generator(device=DEV).manual_seed(11)
    x = torch.randn(2, 32, 2, 128, 128, device=DEV, generator=g).bfloat16()
    flow = torch.randn(2, 2, 2, 128, 128, device=DEV, generator=g) * 0.3
    gout = torch.randn(2, 32, 4, 128, 128, device=DEV, generator=g).bfloat16()
    got = run_warp(x, flow, gout)
    want = torch_ref.warp_with_grads(x.float(), flow, gout.float())
    assert got[0].dtype == torch.bfloat16 and got[1].dtype == torch.bfloat16 and got[2].dtype == torch.float32
    assert float((got[0].float() - want[0]).abs().max()) <= 2e-2
    if bv >= 1:  # variant 0 accumulates the scatter in bf16 atomics; the tiled kernels accumulate in fp32
        assert float((got[1].float() - want[1]).abs().max()) <= 2e-2 * _scale(want[1])
    assert float((got[2] - want[2]).abs().max()) <= 1e-4 * _scale(want[2])


def test_warp_pair_equals_stacked(variants):
    g = torch.Generator(device=DEV).manual_seed(5)
    a = torch.randn(3, 16, 128, 128, device=DEV, generator=g)
    b = torch.randn(3, 16, 128, 128, device=DEV, generator=g)
    flow = torch.randn(3, 2, 2, 128, 128, device=DEV, generator=g)
    gout = torch.randn(3, 16, 4, 128, 128, device=DEV, generator=g)
    ar, br, fr = (t.clone().requires_grad_(True) for t in (a, b, flow))
    out = ops.warp_pair(ar, br, fr)
    out.backward(gout)
    ref = run_warp(torch.stack((a, b), 2), flow, gout)
    assert torch.equal(out.detach(), ref[0])
    assert float((ar.grad - ref[1][:, :, 0]).abs().max()) <= 1e-5 and float((br.grad - ref[1][:, :, 1]).abs().max()) <= 1e-5
    assert float((fr.grad - ref[2]).abs().max()) <= 1e-5 * _scale(ref[2])


@pytest.mark.parametrize("fv", FWD_VARIANTS)
def test_warp_properties_at_full_size(fv, variants):
    """Config 3's per-GPU shape (B=16, C=32, 128x128): linearity in x, exact pass-through slots,
    adjoint identity <warp(x), g> == <x, warp^T(g)> between the forward and backward kernels."""
    _lib.set_option("warp_fwd_variant", fv)
    g = torch.Generator(device=DEV).manual_seed(8)
    x = torch.randn(16, 32, 2, 128, 128, device=DEV, generator=g)
    y = torch.randn(16, 32, 2, 128, 128, device=DEV, generator=g)
    flow = torch.randn(16, 2, 2, 128, 128, device=DEV, generator=g) * 0.5
    gout = torch.randn(16, 32, 4, 128, 128, device=DEV, generator=g)
    with torch.no_grad():
        wx, wy, wxy = (ops.flow_warp(t, flow, (128, 128)) for t in (x, y, x + y))
    assert float((wx + wy - wxy).abs().max()) <= 5e-6
    assert torch.equal(wx[:, :, 0], x[:, :, 0]) and torch.equal(wx[:, :, 3], x[:, :, 1])
    for bv in BWD_VARIANTS:
        _lib.set_option("warp_bwd_variant", bv)
        _, gx, _ = run_warp(x, flow, gout)
        lhs = float((wx.double() * gout.double()).sum())
        rhs = float((x.double() * gx.double()).sum())
        assert abs(lhs - rhs) <= 1e-6 * max(1.0, abs(lhs)) + 1e-2, (lhs, rhs)


@pytest.mark.parametrize("bv", BWD_VARIANTS)
def test_warp_backward_converging_and_far_flows(bv, variants):
    """Backward edge cases of the inverse-gather kernel: many sources landing on one target pixel
    (register list overflow), sources far outside the tile window (atomic side kernel), mixed with
    ordinary sub-pixel flow."""
    _lib.set_option("warp_bwd_variant", bv)
    B, C, H, W = 2, 8, 64, 128
    g = torch.Generator(device=DEV).manual_seed(21)
    x = torch.randn(B, C, 2, H, W, device=DEV, generator=g)
    gout = torch.randn(B, C, 4, H, W, device=DEV, generator=g)
    ws = torch.arange(W, device=DEV, dtype=torch.float32).view(1, 1, 1, W)
    hs = torch.arange(H, device=DEV, dtype=torch.float32).view(1, 1, H, 1)
    flow = torch.randn(B, 2, 2, H, W, device=DEV, generator=g) * 0.3
    # pair 0, frame 0: every pixel samples (almost) the same point -> hundreds of sources per target
    flow[0, 0, 0] = ((W / 2 + 0.37) - ws) * 2 * W / (W - 1)
    flow[0, 1, 0] = ((H / 2 + 0.61) - hs) * 2 * H / (H - 1)
    # pair 1, frame 1: a 20-row / 9-column shear, far beyond HALO / DCAP
    flow[1, 0, 1] += 18.0
    flow[1, 1, 1] += 40.0
    check_warp(run_warp(x, flow, gout), torch_ref.warp_with_grads(x, flow, gout), tol=2e-5)


@pytest.mark.parametrize("bv", [1, 2])
def test_warp_backward_is_deterministic(bv, variants):
    """The tiled backward kernels sum in a fixed order (ATen's atomics do not)."""
    _lib.set_option("warp_bwd_variant", bv)
    g = torch.Generator(device=DEV).manual_seed(4)
    x = torch.randn(4, 32, 2, 128, 128, device=DEV, generator=g)
    flow = torch.randn(4, 2, 2, 128, 128, device=DEV, generator=g) * 0.8
    gout = torch.randn(4, 32, 4, 128, 128, device=DEV, generator=g)
    a, b = run_warp(x, flow, gout), run_warp(x, flow, gout)
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])


def test_warp_argument_errors():
    x = torch.randn(1, 4, 2, 8, 8, device=DEV)
    with pytest.raises(RuntimeError, match="flow must be"):
        ops.flow_warp(x, torch.zeros(1, 2, 2, 8, 9, device=DEV), (8, 8))
    with pytest.raises(RuntimeError, match="must equal"):
        ops.flow_warp(x, torch.zeros(1, 2, 2, 8, 8, device=DEV), (16, 16))
    with pytest.raises(RuntimeError, match="unsupported dtype"):
        ops.flow_warp(x.half(), torch.zeros(1, 2, 2, 8, 8, device=DEV), (8, 8))


# ---------------------------------------------------------------------------------- tlerp + concat
@pytest.mark.parametrize("name", golden_cases("tlerp")[1])
def test_tlerp_against_reference_golden(name):
    z, _ = golden_cases("tlerp")
    has_dec = ("tlerp/%s/dec" % name) in z.files
    g = lambda k: torch.from_numpy(z["tlerp/%s/%s" % (name, k)]).to(DEV)
    skip = g("skip").requires_grad_(True)
    dec = g("dec").requires_grad_(True) if has_dec else None
    cat = ops.tlerp_cat(dec, skip)
    cat.backward(g("gcat"))
    assert float((cat.detach() - g("cat")).abs().max()) <= 1e-6
    assert float((skip.grad - g("gskip")).abs().max()) <= 2e-6
    cd = dec.shape[1] if has_dec else 0
    assert torch.equal(cat[:, cd:, 0], skip[:, :, 0]) and torch.equal(cat[:, cd:, 3], skip[:, :, 1])
    if has_dec:
        assert torch.equal(dec.grad, g("gdec")) and torch.equal(cat[:, :cd].detach(), g("dec"))


# the five scales of both models (Cd, Cs, h) + ragged shapes
TLERP_CASES = [(32, 32, 128), (64, 32, 64), (64, 64, 32), (128, 128, 16), (256, 256, 8), (28, 16, 128), (32, 24, 64),
               (320, 320, 8), (0, 256, 8), (3, 5, 7), (0, 1, 3)]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", TLERP_CASES)
def test_tlerp_against_same_device_reference(case, dtype):
    Cd, Cs, h = case
    B = 3
    g = torch.Generator(device=DEV).manual_seed(Cd + Cs + h)
    skip = torch.randn(B, Cs, 2, h, h, device=DEV, generator=g).to(dtype)
    dec = torch.randn(B, Cd, 4, h, h, device=DEV, generator=g).to(dtype) if Cd else None
    gcat = torch.randn(B, Cd + Cs, 4, h, h, device=DEV, generator=g).to(dtype)
    s1 = skip.clone().requires_grad_(True)
    d1 = dec.clone().requires_grad_(True) if Cd else None
    cat = ops.tlerp_cat(d1, s1)
    cat.backward(gcat)
    want = torch_ref.tlerp_cat_with_grads(None if dec is None else dec.float(), skip.float(), gcat.float())
    tol = 1e-6 if dtype == torch.float32 else 2e-2
    assert float((cat.detach().float() - want[0]).abs().max()) <= tol
    assert float((s1.grad.float() - want[2]).abs().max()) <= (2e-6 if dtype == torch.float32 else 4e-2)
    if Cd:
        assert torch.equal(d1.grad.float(), want[1])
    # un-stacked frames give the same bits
    a, b = skip[:, :, 0].contiguous().requires_grad_(True), skip[:, :, 1].contiguous().requires_grad_(True)
    cat2 = ops.tlerp_pair_cat(dec, a, b)
    cat2.backward(gcat)
    assert torch.equal(cat2.detach(), cat.detach())
    assert torch.equal(a.grad, s1.grad[:, :, 0]) and torch.equal(b.grad, s1.grad[:, :, 1])


def test_empty_batch_launches_nothing():
    """Ragged / empty inputs: a zero-pair batch returns empty tensors of the right shape, like the reference."""
    x = torch.zeros(0, 8, 2, 16, 16, device=DEV, requires_grad=True)
    flow = torch.zeros(0, 2, 2, 16, 16, device=DEV)
    before = _lib.launch_count()
    out = ops.flow_warp(x, flow, (16, 16))
    assert out.shape == (0, 8, 4, 16, 16) and _lib.launch_count() == before
    out.sum().backward()
    assert x.grad.shape == x.shape
    assert ops.tlerp(x.detach()).shape == (0, 8, 4, 16, 16)
    assert ops.tlerp_cat(torch.zeros(0, 3, 4, 16, 16, device=DEV), x.detach()).shape == (0, 11, 4, 16, 16)
    assert ops.warp_pair(x.detach()[:, :, 0], x.detach()[:, :, 1], flow).shape == (0, 8, 4, 16, 16)


def test_large_batch_config3_shape():
    """Config 3's single-GPU shape (B=128, C=32, 128x128: 0.5 GiB in, 1 GiB out): pass-through slots exact,
    warped slots equal to the same-device reference on a sample of pairs."""
    g = torch.Generator(device=DEV).manual_seed(12)
    x = torch.randn(128, 32, 2, 128, 128, device=DEV, generator=g)
    flow = torch.randn(128, 2, 2, 128, 128, device=DEV, generator=g) * 0.4
    with torch.no_grad():
        out = ops.flow_warp(x, flow, (128, 128))
        assert torch.equal(out[:, :, 0], x[:, :, 0]) and torch.equal(out[:, :, 3], x[:, :, 1])
        for b in (0, 63, 127):
            ref = torch_ref.ref_flow_warp(x[b:b + 1], flow[b:b + 1])
            assert float((out[b:b + 1] - ref).abs().max()) <= 1e-6


def test_launch_counter_counts_our_kernels():
    x = torch.randn(1, 4, 2, 16, 16, device=DEV)
    before = _lib.launch_count()
    ops.tlerp(x)
    assert _lib.launch_count() == before + 1


# ---------------------------------------------------------------------------------- channels-last (NDHWC) kernels
CL3 = torch.channels_last_3d


@pytest.mark.parametrize("case", [(4, 16, 128, 128, 0.3), (2, 32, 128, 128, 1.0), (2, 64, 64, 64, 8.0), (1, 8, 37, 52, 2.0),
                                  (1, 256, 32, 32, 0.5), (2, 12, 40, 24, 3.0), (1, 16, 5, 7, 2.5), (2, 32, 33, 128, 1.5),
                                  (1, 4, 16, 1024, 0.8), (1, 512, 16, 16, 1.2)])
@pytest.mark.parametrize("bv", [-1, 0])
def test_warp_ndhwc_against_same_device_reference(case, bv, variants):
    """channels_last_3d in -> channels_last_3d out (vector-gather forward; backward: tile gather + far pass by
    default, vector-atomic scatter as variant 0 and for channel counts the gather does not take)."""
    B, C, H, W, sigma = case
    _lib.set_option("warp_bwd_variant", bv)
    g = torch.Generator(device=DEV).manual_seed(C + H)
    x = torch.randn(B, C, 2, H, W, device=DEV, generator=g).contiguous(memory_format=CL3)
    flow = torch.randn(B, 2, 2, H, W, device=DEV, generator=g) * sigma
    gout = torch.randn(B, C, 4, H, W, device=DEV, generator=g).contiguous(memory_format=CL3)
    xr, fr = x.clone(memory_format=torch.preserve_format).requires_grad_(True), flow.clone().requires_grad_(True)
    before = _lib.launch_count()
    out = ops.flow_warp(xr, fr, (H, W))
    assert out.is_contiguous(memory_format=CL3) and _lib.launch_count() == before + 1
    out.backward(gout)
    assert xr.grad.is_contiguous(memory_format=CL3)
    check_warp((out.detach(), xr.grad, fr.grad), torch_ref.warp_with_grads(x, flow, gout))
    assert torch.equal(out[:, :, 0], x[:, :, 0]) and torch.equal(out[:, :, 3], x[:, :, 1])


def test_warp_ndhwc_gather_backward_is_deterministic_and_matches_scatter(variants):
    """warp_bwd_variant=3 + the caller's workspace: the NDHWC backward is a fixed-order gather (bit-reproducible);
    the default vector-atomic scatter must agree with it."""
    _lib.set_option("warp_bwd_variant", 3)
    g = torch.Generator(device=DEV).manual_seed(31)
    x = torch.randn(3, 16, 2, 128, 128, device=DEV, generator=g).contiguous(memory_format=CL3)
    flow = torch.randn(3, 2, 2, 128, 128, device=DEV, generator=g) * 0.6
    gout = torch.randn(3, 16, 4, 128, 128, device=DEV, generator=g).contiguous(memory_format=CL3)
    before = _lib.launch_count()
    a = run_warp(x, flow, gout)
    assert _lib.launch_count() - before == 1 + 6          # forward + (header, stat, list, apply, 2 early-exit scatter kernels)
    b = run_warp(x, flow, gout)
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    _lib.set_option("warp_bwd_variant", 0)
    before = _lib.launch_count()
    c = run_warp(x, flow, gout)
    assert _lib.launch_count() - before == 1 + 2          # forward + (init, scatter) for one L2-sized chunk
    assert float((a[1] - c[1]).abs().max()) <= 1e-5 and float((a[2] - c[2]).abs().max()) <= 1e-5 * _scale(a[2])


@pytest.mark.parametrize("sigma", [0.3, 1.5])
def test_warp_ndhwc_tile_gather_backward(sigma, variants):
    """Default NDHWC backward: one tile-gather launch (+ the far-tap pass).  Without far taps (sub-pixel flow)
    it is bit-reproducible; it agrees with the vector-atomic scatter kernel."""
    _lib.set_option("warp_bwd_variant", -1)
    g = torch.Generator(device=DEV).manual_seed(77)
    x = torch.randn(3, 32, 2, 128, 128, device=DEV, generator=g).contiguous(memory_format=CL3)
    flow = torch.randn(3, 2, 2, 128, 128, device=DEV, generator=g) * sigma
    gout = torch.randn(3, 32, 4, 128, 128, device=DEV, generator=g).contiguous(memory_format=CL3)
    before = _lib.launch_count()
    a = run_warp(x, flow, gout)
    assert _lib.launch_count() - before == 1 + 2          # forward + (tile gather, far pass)
    b = run_warp(x, flow, gout)
    if sigma < 1.0:
        assert torch.equal(a[1], b[1])
    assert torch.equal(a[2], b[2])
    _lib.set_option("warp_bwd_variant", 0)
    c = run_warp(x, flow, gout)
    assert float((a[2] - c[2]).abs().max()) <= 1e-5 * _scale(a[2])
    assert float((a[1] - c[1]).abs().max()) <= 1e-5
    check_warp(a, torch_ref.warp_with_grads(x, flow, gout))


@pytest.mark.parametrize("case", [(2, 16, 128, 128, 0.3), (1, 64, 40, 72, 0.4), (1, 256, 32, 32, 0.3), (2, 128, 33, 65, 0.5),
                                  (1, 8, 37, 52, 0.3), (1, 32, 5, 7, 0.2)])
def test_warp_ndhwc_tile_gather_tap_prefetch_is_bit_identical(case, variants):
    """Knob ndhwc_bwd_pf: the variant of the tile gather that prefetches the next item's x taps into shared-memory slots
    (auto for C >= 128) performs the same arithmetic in the same order as the plain one: gx and gflow are bit-identical
    (sub-pixel flows: no float atomics involved), and both match the same-device reference."""
    B, C, H, W, sigma = case
    g = torch.Generator(device=DEV).manual_seed(C * 7 + W)
    x = torch.randn(B, C, 2, H, W, device=DEV, generator=g).contiguous(memory_format=CL3)
    flow = (torch.randn(B, 2, 2, H, W, device=DEV, generator=g) * sigma).clamp(-0.9, 0.9)
    gout = torch.randn(B, C, 4, H, W, device=DEV, generator=g).contiguous(memory_format=CL3)
    saved = _lib.get_option("ndhwc_bwd_pf")
    try:
        res = {}
        for pf in (0, 1):
            _lib.set_option("ndhwc_bwd_pf", pf)
            res[pf] = run_warp(x, flow, gout)
    finally:
        _lib.set_option("ndhwc_bwd_pf", saved)
    assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])
    check_warp(res[1], torch_ref.warp_with_grads(x, flow, gout))


@pytest.mark.parametrize("sigma", [0.5, 8.0])
@pytest.mark.parametrize("case", [(2, 16, 64, 64), (1, 64, 40, 72), (1, 256, 16, 32), (2, 32, 128, 128)])
def test_warp_ndhwc_bf16_storage_forward_and_backward(case, sigma):
    """bf16 storage on the layout the modules run (channels_last_3d): the NDHWC kernels themselves (no bounce through
    NCDHW) — shuffle-broadcast forward and the bf16 tile-gather backward with fp32 accumulation + far-tap pass.  Oracle =
    the fp32 reference on the bf16-rounded features (the reference's own bf16 path quantises the grid).
    north_star bar: <= 2e-2 (gradients relative to their max)."""
    B, C, H, W = case
    g = torch.Generator(device=DEV).manual_seed(C + H)
    x = torch.randn(B, C, 2, H, W, device=DEV, generator=g).bfloat16().contiguous(memory_format=CL3)
    flow = torch.randn(B, 2, 2, H, W, device=DEV, generator=g) * sigma
    gout = torch.randn(B, C, 4, H, W, device=DEV, generator=g).bfloat16().contiguous(memory_format=CL3)
    xi, fi = x.clone(memory_format=torch.preserve_format).requires_grad_(True), flow.clone().requires_grad_(True)
    before = _lib.launch_count()
    out = ops.flow_warp(xi, fi, (H, W))
    out.backward(gout)
    assert _lib.launch_count() - before == 3                      # forward, tile gather, far pass: the NDHWC kernels
    assert out.dtype == torch.bfloat16 and out.is_contiguous(memory_format=CL3)
    assert xi.grad.is_contiguous(memory_format=CL3)
    ref = torch_ref.warp_with_grads(x.float(), flow, gout.float())
    assert float((out.float() - ref[0]).abs().max()) <= 2e-2
    assert float((xi.grad.float() - ref[1]).abs().max()) <= 2e-2 * _scale(ref[1])
    assert float((fi.grad - ref[2]).abs().max()) <= 2e-2 * _scale(ref[2])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", [(28, 16, 128), (32, 24, 64), (64, 32, 32), (160, 96, 16), (320, 320, 8), (0, 16, 16), (256, 256, 8)])
def test_tlerp_ndhwc_against_same_device_reference(case, dtype):
    Cd, Cs, h = case
    B = 2
    g = torch.Generator(device=DEV).manual_seed(Cd + Cs + h)
    skip = torch.randn(B, Cs, 2, h, h, device=DEV, generator=g).to(dtype).contiguous(memory_format=CL3)
    dec = torch.randn(B, Cd, 4, h, h, device=DEV, generator=g).to(dtype).contiguous(memory_format=CL3) if Cd else None
    gcat = torch.randn(B, Cd + Cs, 4, h, h, device=DEV, generator=g).to(dtype).contiguous(memory_format=CL3)
    s1 = skip.clone(memory_format=torch.preserve_format).requires_grad_(True)
    d1 = dec.clone(memory_format=torch.preserve_format).requires_grad_(True) if Cd else None
    cat = ops.tlerp_cat(d1, s1)
    vec_ok = Cd % (16 // skip.element_size()) == 0 and Cs % (16 // skip.element_size()) == 0
    assert cat.is_contiguous(memory_format=CL3) == vec_ok or h == 1
    cat.backward(gcat)
    want = torch_ref.tlerp_cat_with_grads(None if dec is None else dec.float(), skip.float(), gcat.float())
    tol = 1e-6 if dtype == torch.float32 else 2e-2
    assert float((cat.detach().float() - want[0]).abs().max()) <= tol
    assert float((s1.grad.float() - want[2]).abs().max()) <= (2e-6 if dtype == torch.float32 else 4e-2)
    if Cd:
        assert torch.equal(d1.grad.float(), want[1])
    # two channels-last frames (the Siamese backbone's outputs) give the same bits
    a = skip[:, :, 0].contiguous(memory_format=torch.channels_last).requires_grad_(True)
    b = skip[:, :, 1].contiguous(memory_format=torch.channels_last).requires_grad_(True)
    cat2 = ops.tlerp_pair_cat(dec, a, b)
    cat2.backward(gcat)
    assert torch.equal(cat2.detach(), cat.detach())
    assert torch.equal(a.grad, s1.grad[:, :, 0]) and torch.equal(b.grad, s1.grad[:, :, 1])


# ---------------------------------------------------------------------------------- N2: semantic tokenizer
_ref_tokens = torch_ref.ref_semantic_tokens      # the reference's own op sequence (models/SMOW_Net.py:176-187)


@pytest.mark.parametrize("case", [(2, 16, 128, 128, 1.0), (3, 32, 128, 128, 0.5), (1, 64, 17, 23, 2.0), (2, 4, 40, 40, 3.0),
                                  (1, 128, 8, 8, 1.0), (2, 8, 96, 64, 0.1)])
def test_semantic_tokens_against_reference_sequence(case):
    B, C, H, W, wscale = case
    g = torch.Generator(device=DEV).manual_seed(B * C + H)
    x = torch.randn(B, C, 4, H, W, device=DEV, generator=g).contiguous(memory_format=CL3)
    weight = torch.randn(8, C, 1, 1, device=DEV, generator=g) * wscale / C ** 0.5
    bias = torch.randn(8, device=DEV, generator=g)
    gt = torch.randn(B, 4, 8, C, device=DEV, generator=g)
    leaves = [t.clone().requires_grad_(True) for t in (x, weight, bias)]
    before = _lib.launch_count()
    tok = ops.semantic_tokens(*leaves)
    tok.backward(gt)
    assert _lib.launch_count() - before == 4
    ref_leaves = [t.double().requires_grad_(True) for t in (x, weight, bias)]
    ref = _ref_tokens(*ref_leaves)
    ref.backward(gt.double())
    assert float((tok.detach() - ref.detach()).abs().max()) <= 1e-5
    for name, a, b in zip(("gx", "gweight", "gbias"), leaves, ref_leaves):
        err = float((a.grad.double() - b.grad).abs().max())
        assert err <= 1e-5 * max(1.0, float(b.grad.abs().max())), (name, err)
    assert leaves[0].grad.is_contiguous(memory_format=CL3)
    # fixed summation order: bit-reproducible
    leaves2 = [t.clone().requires_grad_(True) for t in (x, weight, bias)]
    tok2 = ops.semantic_tokens(*leaves2)
    tok2.backward(gt)
    assert torch.equal(tok2, tok) and all(torch.equal(a.grad, b.grad) for a, b in zip(leaves, leaves2))


def test_semantic_tokens_sharp_attention_and_errors():
    """Large logits (one-hot attention) stay finite; unsupported inputs raise instead of falling back."""
    g = torch.Generator(device=DEV).manual_seed(4)
    x = torch.randn(1, 16, 4, 64, 64, device=DEV, generator=g).contiguous(memory_format=CL3) * 30
    weight = torch.randn(8, 16, 1, 1, device=DEV, generator=g)
    bias = torch.zeros(8, device=DEV)
    tok = ops.semantic_tokens(x, weight, bias)
    ref = _ref_tokens(x.double(), weight.double(), bias.double())
    assert bool(torch.isfinite(tok).all()) and float((tok - ref).abs().max()) <= 1e-3 * float(ref.abs().max())
    with pytest.raises(RuntimeError):
        ops.semantic_tokens(x.cpu(), weight.cpu(), bias.cpu())
    with pytest.raises(RuntimeError):
        ops.semantic_tokens(x[:, :12], weight[:, :12], bias)      # C/4 = 3 is not a power of two


# ------------------------------------------------------------------------------------- rows A1 + N2 fused: warp -> tokens
@pytest.mark.parametrize("case", [(2, 16, 128, 128, 0.7, 1.0), (2, 32, 128, 128, 0.3, 0.5), (1, 16, 24, 40, 8.0, 1.0),
                                  (3, 32, 9, 12, 2.0, 1.0), (1, 16, 33, 65, 0.0, 2.0), (1, 32, 64, 96, 8.0, 1.0)])
def test_warp_tokens_equals_tokenizer_of_the_warped_stack(case):
    """ops.warp_tokens (the stack is produced in shared memory, never in HBM) against the reference's sequence
    OFW.flow_warp -> Transformer_Encoder pooling (models/SMOW_Net.py:612-638 -> :176-187) through the oracle, forward and
    backward (d x, d flow, d conv_a), and bit for bit against the two-launch CUDA path it replaces.  Shapes cover ragged
    chunks (H*W not a multiple of 512), non-power-of-two widths, zero flow and border-clamped sigma = 8 flows."""
    B, C, H, W, sigma, wscale = case
    g = torch.Generator(device=DEV).manual_seed(7 * B + C + H * W)
    x = torch.randn(B, C, 2, H, W, device=DEV, generator=g).contiguous(memory_format=CL3)
    flow = torch.randn(B, 2, 2, H, W, device=DEV, generator=g) * sigma
    weight = torch.randn(8, C, 1, 1, device=DEV, generator=g) * wscale / C ** 0.5
    bias = torch.randn(8, device=DEV, generator=g)
    gt = torch.randn(B, 4, 8, C, device=DEV, generator=g)
    assert ops.warp_tokens_supported(x, weight)
    leaves = [t.clone().requires_grad_(True) for t in (x, flow, weight, bias)]
    before = _lib.launch_count()
    tok = ops.warp_tokens(*leaves)
    assert _lib.launch_count() - before == 2                        # chunk kernel + combine: no warp launch, no stack
    tok.backward(gt)
    # (1) the oracle: the reference's own op sequence on the same device — the warp in fp32 (its coordinate chain is an
    # fp32 contract), the pooling evaluated in fp64 like the stand-alone tokenizer test
    ref_leaves = [x.clone().requires_grad_(True), flow.clone().requires_grad_(True),
                  weight.double().requires_grad_(True), bias.double().requires_grad_(True)]
    ref = _ref_tokens(torch_ref.ref_flow_warp(ref_leaves[0], ref_leaves[1]).double(), *ref_leaves[2:])
    ref.backward(gt.double())
    assert float((tok.detach() - ref.detach()).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))
    for name, a, b in zip(("gx", "gflow", "gweight", "gbias"), leaves, ref_leaves):
        err = float((a.grad.double() - b.grad.double()).abs().max())
        assert err <= 2e-5 * max(1.0, float(b.grad.abs().max())), (name, err)
    # (2) the two-launch CUDA path: the staged rows are the warp kernel's output bit for bit, so everything is identical
    two = [t.clone().requires_grad_(True) for t in (x, flow, weight, bias)]
    tok2 = ops.semantic_tokens(ops.flow_warp(two[0], two[1], (H, W)), two[2], two[3])
    tok2.backward(gt)
    assert torch.equal(tok2, tok)
    for name, a, b in zip(("gx", "gflow", "gweight", "gbias"), leaves, two):
        if name == "gx" and sigma >= 0.7:       # taps farther than a pixel are added by the far pass with L2 reductions: order-free
            assert float((a.grad - b.grad).abs().max()) <= 1e-5 * max(1.0, float(b.grad.abs().max())), name
        else:
            assert torch.equal(a.grad, b.grad), name
    assert leaves[0].grad.is_contiguous(memory_format=CL3)


def test_warp_tokens_argument_errors_and_unsupported_channels():
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(1, 16, 2, 16, 16, device=DEV, generator=g).contiguous(memory_format=CL3)
    flow = torch.zeros(1, 2, 2, 16, 16, device=DEV)
    w, b = torch.randn(8, 16, 1, 1, device=DEV, generator=g), torch.zeros(8, device=DEV)
    with pytest.raises(RuntimeError):
        ops.warp_tokens(x.cpu(), flow.cpu(), w.cpu(), b.cpu())              # no CPU path
    with pytest.raises(RuntimeError):
        ops.warp_tokens(x, flow[:, :, :, :8], w, b)                          # flow of another size
    with pytest.raises(RuntimeError):
        ops.warp_tokens(x, flow, w[:, :8], b)                                # conv_a of another width
    x64 = torch.randn(1, 64, 2, 16, 16, device=DEV, generator=g).contiguous(memory_format=CL3)
    assert not ops.warp_tokens_supported(x64, torch.randn(8, 64, 1, 1, device=DEV))     # the modules then run two launches
    assert not ops.warp_tokens_supported(x.bfloat16(), w)
    # zero flow: frames 1 / 2 resample frames 0 / 3 at their own pixel centres (up to the rounding of the base grid)
    tok = ops.warp_tokens(x, flow, w, b)
    d01, d23 = float((tok[:, 0] - tok[:, 1]).abs().max()), float((tok[:, 2] - tok[:, 3]).abs().max())
    assert d01 <= 1e-4 and d23 <= 1e-4, (d01, d23)


@pytest.mark.parametrize("case", [(2, 16, 128, 128, 0.7), (1, 32, 64, 96, 3.0), (1, 128, 9, 33, 1.0), (2, 4, 31, 17, 5.0),
                                  (1, 256, 16, 24, 1.0), (1, 512, 8, 12, 2.0)])      # > 32 vectors per pixel: channel groups
def test_warp_ndhwc_forward_kernels_agree_bit_for_bit(case, variants):
    """The default NDHWC forward (one coordinate chain per pixel, footprints passed by warp shuffles) and the plain
    per-(pixel, vector) kernel (warp_fwd_variant = 0) run the same arithmetic: identical bits."""
    B, C, H, W, sigma = case
    g = torch.Generator(device=DEV).manual_seed(H * W + C)
    x = torch.randn(B, C, 2, H, W, device=DEV, generator=g).contiguous(memory_format=CL3)
    flow = torch.randn(B, 2, 2, H, W, device=DEV, generator=g) * sigma
    with torch.no_grad():
        _lib.set_option("warp_fwd_variant", -1)
        a = ops.flow_warp(x, flow, (H, W))
        _lib.set_option("warp_fwd_variant", 0)
        b = ops.flow_warp(x, flow, (H, W))
    assert torch.equal(a, b)
    assert float((a - torch_ref.ref_flow_warp(x, flow)).abs().max()) <= 1e-6


# ------------------------------------------------------------------------------------------------ row N1: flow head
@pytest.fixture
def strict_conv():
    """The oracle's F.conv3d must run in full fp32 (cuDNN's TF32 default would make the CHECKER the imprecise side)."""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("case", [(2, 16, 128, 128, 16, 16), (1, 32, 128, 128, 16, 16), (2, 16, 24, 40, 3, 5),
                                  (1, 64, 16, 16, 2, 2), (3, 16, 9, 12, 1, 1), (1, 32, 40, 36, 5, 5)])
def test_flow_head_matches_the_oracle(case, strict_conv):
    """Row N1: ops.flow_head (up-sample + concat + flow_make in one pass, nothing materialised) against the oracle's
    restatement of reference models/SMOW_Net.py:606-608 (interpolate + cat + conv3d) on the same device, strict fp32:
    the flow, d x, d coarse and d weight.  Tolerance 1e-5 of each tensor's scale (the kernel contracts the seg half at low
    resolution, i.e. it re-associates the same fp32 sums).  Ragged shapes cover image borders and coarse grids of 1."""
    from oracle import torch_ref
    from smow_net_b200 import _lib
    B, C, H, W, h, w = case
    g = torch.Generator().manual_seed(B * 100 + C + H)
    x = torch.randn(B, C, 2, H, W, generator=g).to(DEV).contiguous(memory_format=torch.channels_last_3d)
    coarse = torch.randn(B, C, 2, h, w, generator=g).to(DEV)
    weight = (torch.randn(2, 2 * C, 3, 3, 3, generator=g) / (2 * C * 18) ** 0.5).to(DEV)
    gflow = torch.randn(B, 2, 2, H, W, generator=g).to(DEV)
    res = []
    for fn in (ops.flow_head, torch_ref.ref_flow_head):
        # the oracle arm gets plain contiguous tensors (what the reference itself runs on): cuDNN's strict-fp32 conv3d has
        # no engine for some channels-last ragged shapes — a property of the CHECKER's library, not of the kernel
        fmt = torch.preserve_format if fn is ops.flow_head else torch.contiguous_format
        xi, ci, wi = (t.detach().clone(memory_format=fmt).requires_grad_(True) for t in (x, coarse, weight))
        before = _lib.launch_count()
        flow = fn(xi, ci, wi)
        flow.backward(gflow)
        res.append((flow.detach(), xi.grad, ci.grad, wi.grad, _lib.launch_count() - before))
    assert res[0][4] == 7 and res[1][4] == 0                       # pack + fwd, pack + d x + d Z + d W + reduce
    assert res[0][0].is_contiguous() and res[0][0].shape == (B, 2, 2, H, W)
    for a, b, name in zip(res[0][:4], res[1][:4], ("flow", "gx", "gcoarse", "gweight")):
        assert a.shape == b.shape
        err = float((a - b).abs().max()) / max(1.0, float(b.abs().max()))
        assert err <= 1e-5, (name, err)


def test_flow_head_is_deterministic_and_refuses_cpu():
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 16, 2, 32, 32, generator=g).to(DEV).contiguous(memory_format=torch.channels_last_3d)
    coarse, weight = torch.randn(2, 16, 2, 4, 4, generator=g).to(DEV), torch.randn(2, 32, 3, 3, 3, generator=g).to(DEV)
    outs = []
    for _ in range(2):
        xi, ci, wi = (t.detach().clone(memory_format=torch.preserve_format).requires_grad_(True) for t in (x, coarse, weight))
        f = ops.flow_head(xi, ci, wi)
        f.sum().backward()
        outs.append((f.detach(), xi.grad, ci.grad, wi.grad))
    assert all(torch.equal(a, b) for a, b in zip(*outs))
    with pytest.raises(RuntimeError):
        ops.flow_head(x.cpu(), coarse.cpu(), weight.cpu())


# ------------------------------------------------------------------------------------------------ A3+A4 with the LeakyReLU folded in
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", [(2, 28, 16, 40, 36, True), (3, 32, 24, 9, 7, False), (1, 320, 96, 8, 8, False), (2, 64, 32, 64, 64, True)])
def test_act_tlerp_cat_matches_the_reference_sequence(case, dtype):
    """ops.act_tlerp_cat / act_tlerp_pair_cat = the reference's LeakyReLU(0.2) (models/SMOW_Net.py:137), trilinear 2 -> 4
    (:64-73) and torch.cat (:78-94) in one launch, against that very ATen sequence on the same device: forward bit-exact
    in fp32 for the activation half and frames 0 / 3, <= 1e-6 for the lerped frames; gradients <= 1e-6 (fp32) / 2e-2 (bf16)."""
    import torch.nn.functional as F
    B, Cd, Cs, h, w, stacked = case
    if dtype == torch.bfloat16 and (Cd % 8 or Cs % 8):
        pytest.skip("bf16 vectors hold 8 channels: the modules fall back to the un-fused pair for such channel counts")
    g = torch.Generator().manual_seed(Cd + Cs)
    z = torch.randn(B, Cd, 4, h, w, generator=g).to(DEV, dtype).contiguous(memory_format=torch.channels_last_3d)
    skip = torch.randn(B, Cs, 2, h, w, generator=g).to(DEV, dtype).contiguous(memory_format=torch.channels_last_3d)
    gcat = torch.randn(B, Cd + Cs, 4, h, w, generator=g).to(DEV, dtype)
    zi, si = z.clone(memory_format=torch.preserve_format).requires_grad_(True), skip.clone(memory_format=torch.preserve_format).requires_grad_(True)
    before = _lib.launch_count()
    if stacked:
        cat = ops.act_tlerp_cat(zi, si, 0.2)
    else:
        cat = ops.act_tlerp_pair_cat(zi, si[:, :, 0], si[:, :, 1], 0.2)
    cat.backward(gcat)
    assert _lib.launch_count() - before == 2                          # ONE launch forward, ONE backward
    zr, sr = z.float().clone().requires_grad_(True), skip.float().clone().requires_grad_(True)
    want = torch_ref.ref_tlerp_cat(F.leaky_relu(zr, 0.2), sr)
    want.backward(gcat.float())
    tol = 1e-6 if dtype == torch.float32 else 2e-2
    if dtype == torch.float32:
        assert torch.equal(cat[:, :Cd], want[:, :Cd]) and torch.equal(cat[:, Cd:, 0], want[:, Cd:, 0]) and torch.equal(cat[:, Cd:, 3], want[:, Cd:, 3])
    assert float((cat.float() - want).abs().max()) <= tol
    assert float((zi.grad.float() - zr.grad).abs().max()) <= tol * _scale(zr.grad)
    assert float((si.grad.float() - sr.grad).abs().max()) <= tol * _scale(sr.grad)
    assert cat.is_contiguous(memory_format=torch.channels_last_3d)
