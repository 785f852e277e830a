"""GPU: out-of-bounds canaries.  compute-sanitizer is closed on the B200 pool, so every output buffer of a
direct C-ABI call is surrounded by guard regions that must stay untouched, for every kernel variant and for
shapes that exercise partial tiles (H not a multiple of the band height, first/last bands, tiny images)."""
import pytest
import torch

from smow_net_b200 import _lib, ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GUARD = 4096          # floats on each side
MAGIC = 1234.5


def guarded(numel):
    buf = torch.full((numel + 2 * GUARD,), MAGIC, device=DEV)
    return buf, buf[GUARD:GUARD + numel]


def intact(buf, numel):
    return bool((buf[:GUARD] == MAGIC).all()) and bool((buf[GUARD + numel:] == MAGIC).all())


SHAPES = [(2, 16, 128, 128), (1, 8, 130, 128), (3, 4, 7, 64), (2, 32, 64, 64), (1, 4, 33, 256), (1, 12, 100, 72)]


@pytest.mark.parametrize("bv", [0, 1, 2])
@pytest.mark.parametrize("fv", [0, 1, 2])
@pytest.mark.parametrize("shape", SHAPES)
def test_warp_kernels_stay_inside_their_buffers(shape, fv, bv):
    B, C, H, W = shape
    lib = _lib.load()
    saved = {k: _lib.get_option(k) for k in ("warp_fwd_variant", "warp_bwd_variant")}
    _lib.set_option("warp_fwd_variant", fv)
    _lib.set_option("warp_bwd_variant", bv)
    try:
        g = torch.Generator(device=DEV).manual_seed(1)
        x = torch.randn(B, C, 2, H, W, device=DEV, generator=g)
        flow = torch.randn(B, 2, 2, H, W, device=DEV, generator=g) * 3.0
        gout = torch.randn(B, C, 4, H, W, device=DEV, generator=g)
        xs, ys = ops.base_grid(W, x.device), ops.base_grid(H, x.device)
        st = torch.cuda.current_stream().cuda_stream
        n_out, n_gx, n_gf = B * C * 4 * H * W, B * C * 2 * H * W, B * 2 * 2 * H * W
        ob, out = guarded(n_out)
        _lib.check(lib.smow_warp_stack_fwd(x.data_ptr(), flow.data_ptr(), xs.data_ptr(), ys.data_ptr(), out.data_ptr(),
                                           B, C, H, W, 0, 0, st), "fwd")
        gb, gx = guarded(n_gx)
        fb, gf = guarded(n_gf)
        _lib.check(lib.smow_warp_stack_bwd(gout.data_ptr(), x.data_ptr(), flow.data_ptr(), xs.data_ptr(), ys.data_ptr(),
                                           gx.data_ptr(), gf.data_ptr(), B, C, H, W, 0, 0, None, 0, st), "bwd")
        torch.cuda.synchronize()
        assert intact(ob, n_out) and intact(gb, n_gx) and intact(fb, n_gf)
        assert bool((out != MAGIC).all()) and bool((gx != MAGIC).all()) and bool((gf != MAGIC).all())   # fully written
        # and the values agree with the autograd path (which the parity tests cover)
        xr, fr = x.clone().requires_grad_(True), flow.clone().requires_grad_(True)
        ref = ops.flow_warp(xr, fr, (H, W))
        assert torch.equal(ref.detach().flatten(), out)
    finally:
        for k, v in saved.items():
            _lib.set_option(k, v)


@pytest.mark.parametrize("case", [(2, 5, 6, 16), (1, 0, 3, 64), (3, 28, 16, 128 * 128), (2, 3, 5, 7)])
def test_tlerp_kernels_stay_inside_their_buffers(case):
    B, Cd, Cs, hw = case
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(2)
    skip = torch.randn(B, Cs, 2, hw, device=DEV, generator=g)
    dec = torch.randn(B, max(Cd, 1), 4, hw, device=DEV, generator=g)
    st = torch.cuda.current_stream().cuda_stream
    n_cat, n_gs = B * (Cd + Cs) * 4 * hw, B * Cs * 2 * hw
    cb, cat = guarded(n_cat)
    _lib.check(lib.smow_tlerp_cat_fwd(dec.data_ptr() if Cd else None, skip.data_ptr(), cat.data_ptr(), B, Cd, Cs, hw, 0, 0, st), "fwd")
    gcat = torch.randn(n_cat, device=DEV, generator=g)
    sb, gs = guarded(n_gs)
    _lib.check(lib.smow_tlerp_cat_bwd(gcat.data_ptr(), gs.data_ptr(), B, Cd, Cs, hw, 0, 0, st), "bwd")
    torch.cuda.synchronize()
    assert intact(cb, n_cat) and intact(sb, n_gs)
    assert bool((cat != MAGIC).all()) and bool((gs != MAGIC).all())


@pytest.mark.parametrize("shape", [(2, 16, 128, 128), (1, 32, 9, 12), (3, 16, 33, 65), (1, 32, 64, 96), (2, 16, 24, 40)])
def test_fused_warp_tokenizer_and_bn_backward_stay_inside_their_buffers(shape):
    """Direct C-ABI calls of the fused warp -> tokens pass (ragged last chunk, non-power-of-two widths, sigma = 3 flows that
    push taps to the image border) and of the row-wise BatchNorm + lerp backward (channel splits whose rows are not a
    multiple of 128 bytes): every output sits between guard regions."""
    B, C, H, W = shape
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(5)
    st = torch.cuda.current_stream().cuda_stream
    x = torch.randn(B, C, 2, H, W, device=DEV, generator=g).contiguous(memory_format=torch.channels_last_3d)
    flow = torch.randn(B, 2, 2, H, W, device=DEV, generator=g) * 3.0
    wa, ba = torch.randn(8, C, device=DEV, generator=g) / 4, torch.randn(8, device=DEV, generator=g)
    xs, ys = ops.base_grid(W, x.device), ops.base_grid(H, x.device)
    n_ws = int(lib.smow_tokenizer_workspace_bytes(B, C, H * W))
    wsb, ws = guarded(n_ws // 4)
    tb, tokens = guarded(B * 4 * 8 * C)
    sb, stats = guarded(B * 4 * 16)
    _lib.check(lib.smow_warp_tokenizer_fwd(x.data_ptr(), flow.data_ptr(), xs.data_ptr(), ys.data_ptr(), wa.data_ptr(), ba.data_ptr(),
                                           tokens.data_ptr(), stats.data_ptr(), B, C, H, W, 0, 1, ws.data_ptr(), n_ws, st), "fwd")
    gt = torch.randn(B, 4, 8, C, device=DEV, generator=g)
    n_gs = B * C * 4 * H * W
    gb, gstack = guarded(n_gs)
    wb, gwa = guarded(8 * C)
    bb, gba = guarded(8)
    _lib.check(lib.smow_warp_tokenizer_bwd(gt.data_ptr(), x.data_ptr(), flow.data_ptr(), xs.data_ptr(), ys.data_ptr(), wa.data_ptr(),
                                           ba.data_ptr(), tokens.data_ptr(), stats.data_ptr(), gstack.data_ptr(), gwa.data_ptr(),
                                           gba.data_ptr(), B, C, H, W, 0, 1, ws.data_ptr(), n_ws, st), "bwd")
    torch.cuda.synchronize()
    assert intact(wsb, n_ws // 4) and intact(tb, tokens.numel()) and intact(sb, stats.numel())
    assert intact(gb, n_gs) and intact(wb, 8 * C) and intact(bb, 8)
    assert bool((tokens != MAGIC).all()) and bool((gstack != MAGIC).all()) and bool((gwa != MAGIC).all())
    # same values as the autograd path (which the parity tests hold against the oracle)
    ref = ops.warp_tokens(x, flow, wa.view(8, C, 1, 1), ba)
    assert torch.equal(ref.flatten(), tokens)
    # ---- row-wise BatchNorm + LeakyReLU + lerp backward: 12 + 16 / 28 + 16 channels (rows of 112 / 176 bytes), hw = H * W
    Cd, Cs, hw = (12 if C == 16 else 28), 16, H * W
    gcat = torch.randn(B * 4 * hw * (Cd + Cs), device=DEV, generator=g)
    y = torch.randn(B * 4 * hw * Cd, device=DEV, generator=g)
    bn = torch.rand(6, Cd, device=DEV, generator=g) + 0.5
    yb, gy = guarded(B * 4 * hw * Cd)
    kb, gskip = guarded(B * 2 * hw * Cs)
    _lib.check(lib.smow_bn_act_tlerp_cat_bwd(gcat.data_ptr(), y.data_ptr(), bn.data_ptr(), gy.data_ptr(), gskip.data_ptr(),
                                             gskip.data_ptr() + hw * Cs * 4, B, Cd, Cs, hw, 2 * Cs * hw, 0.2, st), "bn bwd")
    torch.cuda.synchronize()
    assert intact(yb, gy.numel()) and intact(kb, gskip.numel())
    assert bool((gy != MAGIC).all()) and bool((gskip != MAGIC).all())
    # the row-wise kernel and the split decoder / skip kernel compute the same thing
    _lib.set_option("bn_bwd_rows", 0)
    try:
        gy0, gs0 = torch.empty_like(gy), torch.empty_like(gskip)
        _lib.check(lib.smow_bn_act_tlerp_cat_bwd(gcat.data_ptr(), y.data_ptr(), bn.data_ptr(), gy0.data_ptr(), gs0.data_ptr(),
                                                 gs0.data_ptr() + hw * Cs * 4, B, Cd, Cs, hw, 2 * Cs * hw, 0.2, st), "bn bwd (split)")
    finally:
        _lib.set_option("bn_bwd_rows", 1)
    torch.cuda.synchronize()
    assert torch.equal(gy0, gy) and torch.equal(gs0, gskip)
