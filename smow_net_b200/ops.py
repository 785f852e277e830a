"""Autograd operators of the SMOW-Net alignment/fusion hot path (CUDA only).

Python mirror of the reference's operator interface for this path:

* ``flow_warp(input, flow, size)``     - ``OFW.flow_warp`` (reference models/SMOW_Net.py:612-638)
* ``warp_pair(x_t1, x_t2, flow)``      - same, frames not yet stacked (models/SMOW_Net_LW.py:38-40,58)
* ``tlerp_cat(dec, skip)``             - ``F.interpolate(skip, (4,h,w), 'trilinear', True)`` +
  ``torch.cat([dec, skip_up], 1)``     (models/SMOW_Net.py:64-73,78-94)
* ``tlerp(skip)`` / ``tlerp_pair_cat`` - stand-alone / un-stacked forms of the same

Every operator calls libsmow_b200.so through the C ABI of include/smow_b200.h.  There is
no CPU implementation and no PyTorch fallback: CPU tensors raise.
"""
import contextlib

import torch

from . import _lib

_DTYPES = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}


# ----------------------------------------------------------------------------- helpers
def _require_cuda(*tensors):
    """Every operand is a CUDA tensor and all of them live on ONE device (the C ABI takes raw pointers and launches on
    the current device: it cannot catch either mistake)."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "smow_net_b200 operators run on CUDA tensors only (hand-written sm_100a kernels, "
                "no CPU fallback); got a %s tensor" % t.device)
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("smow_net_b200: operands on different devices (%s and %s)" % (dev, t.device))


def _dtype_code(t):
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise RuntimeError("smow_net_b200: unsupported dtype %s (fp32 and bf16 only)" % t.dtype)


def _is_channels_last(t):
    fmt = torch.channels_last_3d if t.dim() == 5 else torch.channels_last
    return (not t.is_contiguous()) and t.is_contiguous(memory_format=fmt)


def _vec(t):
    return 16 // t.element_size()          # channels per 16-byte vector


def _layout5(t, *channel_counts, ndhwc_ok=True):
    """(dense tensor, layout code).  A channels_last / channels_last_3d tensor keeps its order (NDHWC kernels)
    when every channel count is a multiple of the 16-byte vector; anything else becomes contiguous NCDHW."""
    if ndhwc_ok and _is_channels_last(t) and all(c % _vec(t) == 0 for c in channel_counts):
        return t, _lib.NDHWC
    return t.contiguous(), _lib.NCDHW


def _ndhwc_warp_ok(t, C):
    """The NDHWC warp kernels take fp32 (any C % 4 == 0) and bf16 storage with C / 8 a power of two (the tile gather's
    lanes-per-pixel split); other bf16 channel counts go through the NCDHW kernels."""
    if t.dtype == torch.float32:
        return True
    q = C // 8
    return t.dtype == torch.bfloat16 and C % 8 == 0 and q > 0 and (q & (q - 1)) == 0


def _as_layout(t, layout):
    if layout == _lib.NCDHW:
        return t.contiguous()
    fmt = torch.channels_last_3d if t.dim() == 5 else torch.channels_last
    return t.contiguous(memory_format=fmt)


def _empty(shape, like, layout):
    fmt = torch.contiguous_format
    if layout == _lib.NDHWC:
        fmt = torch.channels_last_3d if len(shape) == 5 else torch.channels_last
    return torch.empty(shape, dtype=like.dtype, device=like.device, memory_format=fmt)


def _stream():
    return torch.cuda.current_stream().cuda_stream


_GRID_CACHE = {}


def base_grid(n, device):
    """fp32 table torch.linspace(-1, 1, n) built on the CPU exactly as the reference does
    (models/SMOW_Net.py:617-618) and cached on the device, so the coordinate chain is
    reproduced bit for bit without the reference's per-call host-to-device copy."""
    key = (int(n), device.index if device.index is not None else torch.cuda.current_device())
    t = _GRID_CACHE.get(key)
    if t is None:
        t = torch.linspace(-1.0, 1.0, int(n)).to(device)
        _GRID_CACHE[key] = t
    return t


# ------------------------------------------------------------------- per-launch timing hook
class KernelTimer:
    """Collects CUDA-event timings of every C-ABI call made while active (bench.py roofline)."""

    def __init__(self):
        self.records = []  # (name, algorithmic_bytes, start_event, end_event, meta)

    def calls(self):
        """[(name, meta dict)] of every recorded C-ABI call, in launch order (smow_net_b200.probe replays them)."""
        return [(r[0], r[4]) for r in self.records]

    def summary(self, by_shape=False):
        """Per operator totals; by_shape=True keeps launches of different sizes apart (key 'name@<bytes>')."""
        out = {}
        for name, nbytes, e0, e1, _meta_ in self.records:
            ms = e0.elapsed_time(e1)
            s = out.setdefault("%s@%d" % (name, nbytes) if by_shape else name, {"calls": 0, "ms": 0.0, "bytes": 0})
            s["calls"] += 1
            s["ms"] += ms
            s["bytes"] += nbytes
        for s in out.values():
            s["gbps"] = s["bytes"] / (s["ms"] * 1e-3) / 1e9 if s["ms"] > 0 else 0.0
        return out


_TIMER = None


@contextlib.contextmanager
def kernel_timer():
    global _TIMER
    prev, _TIMER = _TIMER, KernelTimer()
    try:
        yield _TIMER
    finally:
        _TIMER = prev


_NEXT_META = None


def _meta(**kw):
    """Shape / layout of the next C-ABI call, kept only while a kernel_timer is active."""
    global _NEXT_META
    if _TIMER is not None:
        _NEXT_META = kw


def _call(name, nbytes, fn, *args):
    global _NEXT_META
    if _TIMER is None:
        _lib.check(fn(*args), name)
        return
    meta, _NEXT_META = _NEXT_META, None
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(fn(*args), name)
    e1.record()
    _TIMER.records.append((name, nbytes, e0, e1, meta))


# algorithmic bytes per launch; s = bytes per feature element, flow is fp32
def warp_fwd_bytes(B, C, H, W, s):
    return B * (6 * C * H * W * s + 16 * H * W)


def warp_bwd_bytes(B, C, H, W, s):
    return B * (8 * C * H * W * s + 32 * H * W)


def tlerp_fwd_bytes(B, Cd, Cs, hw, s):
    """read 2*Cs*hw*s, write 4*Cs*hw*s.  The copy of the decoder half is NOT algorithmic."""
    return B * 6 * Cs * hw * s


def tlerp_fwd_overhead_bytes(B, Cd, hw, s):
    """Extra traffic of the same launch when it also copies the decoder half into the concat buffer (read + write)."""
    return B * 8 * Cd * hw * s


def tlerp_bwd_bytes(B, Cs, hw, s):
    return B * 6 * Cs * hw * s


_FLAG_WS = {}


def _bwd_workspace(lib, like, layout, B, H, W):
    """Caller-owned scratch of the NDHWC fp32 backward (the C ABI never allocates): the gather-list variant (3) needs
    per-pixel lists; the default tile gather only a 64-byte word block for its far-tap stamp — zero-initialised once
    and cached per (device, stream), because two launches in flight must not share a stamp."""
    if layout != _lib.NDHWC:
        return None, 0
    if like.dtype == torch.float32 and _lib.get_option("warp_bwd_variant") == 3:
        n = int(lib.smow_warp_bwd_workspace_bytes(B, H, W))
        return torch.empty(n, dtype=torch.uint8, device=like.device), n
    if torch.cuda.is_current_stream_capturing():
        # inside a CUDA-graph capture the block must belong to the graph's private pool (a cached tensor from outside
        # could be freed under the graph): one 64-byte allocation + memset node per captured backward
        return torch.zeros(64, dtype=torch.uint8, device=like.device), 64
    key = (like.device.index, torch.cuda.current_stream(like.device).cuda_stream)
    ws = _FLAG_WS.get(key)
    if ws is None:
        ws = _FLAG_WS[key] = torch.zeros(64, dtype=torch.uint8, device=like.device)
    return ws, 64


# ----------------------------------------------------------------------------- A1: warp + stack
class _WarpStack(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, flow):
        _require_cuda(x, flow)
        if x.dim() != 5 or x.shape[2] != 2:
            raise RuntimeError("flow_warp: input must be (B,C,2,H,W), got %s" % (tuple(x.shape),))
        B, C, _, H, W = x.shape
        if tuple(flow.shape) != (B, 2, 2, H, W):
            raise RuntimeError("flow_warp: flow must be (B,2,2,H,W)=%s, got %s" % ((B, 2, 2, H, W), tuple(flow.shape)))
        x, layout = _layout5(x, C, ndhwc_ok=_ndhwc_warp_ok(x, C))
        ctx.flow_dtype = flow.dtype
        flow = flow.float().contiguous()
        out = _empty((B, C, 4, H, W), x, layout)
        xs, ys = base_grid(W, x.device), base_grid(H, x.device)
        lib = _lib.load()
        with torch.cuda.device_of(x):
            _meta(B=B, C=C, H=H, W=W, dtype=_dtype_code(x), layout=layout, pair=0)
            _call("warp_stack_fwd", warp_fwd_bytes(B, C, H, W, x.element_size()), lib.smow_warp_stack_fwd,
                  x.data_ptr(), flow.data_ptr(), xs.data_ptr(), ys.data_ptr(), out.data_ptr(),
                  B, C, H, W, _dtype_code(x), layout, _stream())
        ctx.save_for_backward(x, flow)
        ctx.layout = layout
        return out

    @staticmethod
    def backward(ctx, gout):
        x, flow = ctx.saved_tensors
        B, C, _, H, W = x.shape
        gout = _as_layout(gout, ctx.layout)
        gx = _empty(tuple(x.shape), x, ctx.layout)
        gflow = torch.empty_like(flow)
        xs, ys = base_grid(W, x.device), base_grid(H, x.device)
        lib = _lib.load()
        ws, ws_bytes = _bwd_workspace(lib, x, ctx.layout, B, H, W)
        with torch.cuda.device_of(x):
            _meta(B=B, C=C, H=H, W=W, dtype=_dtype_code(x), layout=ctx.layout, pair=0)
            _call("warp_stack_bwd", warp_bwd_bytes(B, C, H, W, x.element_size()), lib.smow_warp_stack_bwd,
                  gout.data_ptr(), x.data_ptr(), flow.data_ptr(), xs.data_ptr(), ys.data_ptr(),
                  gx.data_ptr(), gflow.data_ptr(), B, C, H, W, _dtype_code(x), ctx.layout,
                  ws.data_ptr() if ws is not None else None, ws_bytes, _stream())
        return gx, gflow.to(ctx.flow_dtype)          # the kernels' flow gradient is fp32; a bf16 flow gets it rounded


class _WarpPair(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x1, x2, flow):
        _require_cuda(x1, x2, flow)
        if x1.dim() != 4 or x1.shape != x2.shape or x1.dtype != x2.dtype:
            raise RuntimeError("warp_pair: x_t1/x_t2 must be matching (B,C,H,W) tensors")
        B, C, H, W = x1.shape
        if tuple(flow.shape) != (B, 2, 2, H, W):
            raise RuntimeError("warp_pair: flow must be (B,2,2,H,W)")
        x1, layout = _layout5(x1, C, ndhwc_ok=_ndhwc_warp_ok(x1, C))
        x2 = _as_layout(x2, layout)
        ctx.flow_dtype = flow.dtype
        flow = flow.float().contiguous()
        out = _empty((B, C, 4, H, W), x1, layout)
        xs, ys = base_grid(W, x1.device), base_grid(H, x1.device)
        lib = _lib.load()
        with torch.cuda.device_of(x1):
            _meta(B=B, C=C, H=H, W=W, dtype=_dtype_code(x1), layout=layout, pair=1)
            _call("warp_stack_fwd", warp_fwd_bytes(B, C, H, W, x1.element_size()), lib.smow_warp_pair_fwd,
                  x1.data_ptr(), x2.data_ptr(), flow.data_ptr(), xs.data_ptr(), ys.data_ptr(), out.data_ptr(),
                  B, C, H, W, _dtype_code(x1), layout, _stream())
        ctx.save_for_backward(x1, x2, flow)
        ctx.layout = layout
        return out

    @staticmethod
    def backward(ctx, gout):
        x1, x2, flow = ctx.saved_tensors
        B, C, H, W = x1.shape
        gout = _as_layout(gout, ctx.layout)
        g1 = _empty(tuple(x1.shape), x1, ctx.layout)
        g2 = _empty(tuple(x1.shape), x1, ctx.layout)
        gflow = torch.empty_like(flow)
        xs, ys = base_grid(W, x1.device), base_grid(H, x1.device)
        lib = _lib.load()
        with torch.cuda.device_of(x1):
            ws, ws_bytes = _bwd_workspace(lib, x1, ctx.layout, B, H, W)
            _meta(B=B, C=C, H=H, W=W, dtype=_dtype_code(x1), layout=ctx.layout, pair=1)
            _call("warp_stack_bwd", warp_bwd_bytes(B, C, H, W, x1.element_size()), lib.smow_warp_pair_bwd,
                  gout.data_ptr(), x1.data_ptr(), x2.data_ptr(), flow.data_ptr(), xs.data_ptr(), ys.data_ptr(),
                  g1.data_ptr(), g2.data_ptr(), gflow.data_ptr(), B, C, H, W, _dtype_code(x1), ctx.layout,
                  ws.data_ptr() if ws is not None else None, ws_bytes, _stream())
        return g1, g2, gflow.to(ctx.flow_dtype)


def flow_warp(input, flow, size=None):
    """Drop-in for ``OFW.flow_warp(input, flow, size)``: (B,C,2,H,W),(B,2,2,H,W) -> (B,C,4,H,W).

    ``size`` is accepted for signature compatibility; like the reference it must equal the
    spatial size of ``input`` (the reference's cat at :636 fails otherwise)."""
    if size is not None and tuple(size) != tuple(input.shape[3:]):
        raise RuntimeError("flow_warp: size %s must equal the input's spatial size %s"
                           % (tuple(size), tuple(input.shape[3:])))
    if input.shape[0] == 0:      # empty batch: nothing to launch (the reference returns an empty stack too)
        _require_cuda(input, flow)
        return input.new_zeros((0, input.shape[1], 4) + tuple(input.shape[3:])) + 0 * (input.sum() + flow.sum())
    return _WarpStack.apply(input, flow)


def warp_pair(x_t1, x_t2, flow):
    """flow_warp on two un-stacked (B,C,H,W) frames -> (B,C,4,H,W)."""
    if x_t1.shape[0] == 0:
        return flow_warp(torch.stack((x_t1, x_t2), 2), flow)
    return _WarpPair.apply(x_t1, x_t2, flow)


# ----------------------------------------------------------------------------- A3+A4: tlerp + concat
class _TLerpCat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dec, skip):
        _require_cuda(dec, skip)
        if skip.dim() != 5 or skip.shape[2] != 2:
            raise RuntimeError("tlerp_cat: skip must be (B,Cs,2,h,w), got %s" % (tuple(skip.shape),))
        B, Cs, _, h, w = skip.shape
        Cd = 0
        if dec is not None:
            if dec.dim() != 5 or dec.shape[0] != B or dec.shape[2] != 4 or tuple(dec.shape[3:]) != (h, w):
                raise RuntimeError("tlerp_cat: dec must be (B,Cd,4,h,w) matching skip, got %s" % (tuple(dec.shape),))
            if dec.dtype != skip.dtype:
                raise RuntimeError("tlerp_cat: dec and skip dtypes differ")
            Cd = dec.shape[1]
        # the concat follows the decoder tensor's order (the larger operand), else the skip's
        skip, layout = _layout5(skip if dec is None or not _is_channels_last(dec) else
                                skip.contiguous(memory_format=torch.channels_last_3d), Cd, Cs)
        if dec is not None:
            dec = _as_layout(dec, layout)
        cat = _empty((B, Cd + Cs, 4, h, w), skip, layout)
        lib = _lib.load()
        with torch.cuda.device_of(skip):
            _meta(B=B, Cd=Cd, Cs=Cs, hw=h * w, dtype=_dtype_code(skip), layout=layout, pair=0)
            _call("tlerp_cat_fwd", tlerp_fwd_bytes(B, Cd, Cs, h * w, skip.element_size()), lib.smow_tlerp_cat_fwd,
                  dec.data_ptr() if dec is not None else None, skip.data_ptr(), cat.data_ptr(),
                  B, Cd, Cs, h * w, _dtype_code(skip), layout, _stream())
        ctx.dims = (B, Cd, Cs, h, w)
        ctx.layout = layout
        return cat

    @staticmethod
    def backward(ctx, gcat):
        B, Cd, Cs, h, w = ctx.dims
        gcat = _as_layout(gcat, ctx.layout)
        gskip = _empty((B, Cs, 2, h, w), gcat, ctx.layout)
        lib = _lib.load()
        with torch.cuda.device_of(gcat):
            _meta(B=B, Cd=Cd, Cs=Cs, hw=h * w, dtype=_dtype_code(gcat), layout=ctx.layout, pair=0)
            _call("tlerp_cat_bwd", tlerp_bwd_bytes(B, Cs, h * w, gcat.element_size()), lib.smow_tlerp_cat_bwd,
                  gcat.data_ptr(), gskip.data_ptr(), B, Cd, Cs, h * w, _dtype_code(gcat), ctx.layout, _stream())
        gdec = gcat[:, :Cd] if Cd > 0 else None  # strided view, as torch.cat's backward returns
        return gdec, gskip


class _TLerpPairCat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dec, a, b):
        _require_cuda(dec, a, b)
        if a.dim() != 4 or a.shape != b.shape or a.dtype != b.dtype:
            raise RuntimeError("tlerp_pair_cat: frames must be matching (B,Cs,h,w) tensors")
        B, Cs, h, w = a.shape
        Cd = 0
        if dec is not None:
            if dec.dim() != 5 or dec.shape[0] != B or dec.shape[2] != 4 or tuple(dec.shape[3:]) != (h, w):
                raise RuntimeError("tlerp_pair_cat: dec must be (B,Cd,4,h,w) matching the frames")
            if dec.dtype != a.dtype:
                raise RuntimeError("tlerp_pair_cat: dec and frame dtypes differ (%s vs %s)" % (dec.dtype, a.dtype))
            Cd = dec.shape[1]
        a, layout = _layout5(a if dec is None or not _is_channels_last(dec) else
                             a.contiguous(memory_format=torch.channels_last), Cd, Cs)
        b = _as_layout(b, layout)
        if dec is not None:
            dec = _as_layout(dec, layout)
        cat = _empty((B, Cd + Cs, 4, h, w), a, layout)
        lib = _lib.load()
        with torch.cuda.device_of(a):
            _meta(B=B, Cd=Cd, Cs=Cs, hw=h * w, dtype=_dtype_code(a), layout=layout, pair=1)
            _call("tlerp_cat_fwd", tlerp_fwd_bytes(B, Cd, Cs, h * w, a.element_size()), lib.smow_tlerp_pair_cat_fwd,
                  dec.data_ptr() if dec is not None else None, a.data_ptr(), b.data_ptr(), cat.data_ptr(),
                  B, Cd, Cs, h * w, _dtype_code(a), layout, _stream())
        ctx.dims = (B, Cd, Cs, h, w)
        ctx.layout = layout
        return cat

    @staticmethod
    def backward(ctx, gcat):
        B, Cd, Cs, h, w = ctx.dims
        gcat = _as_layout(gcat, ctx.layout)
        ga = _empty((B, Cs, h, w), gcat, ctx.layout)
        gb = _empty((B, Cs, h, w), gcat, ctx.layout)
        lib = _lib.load()
        with torch.cuda.device_of(gcat):
            _meta(B=B, Cd=Cd, Cs=Cs, hw=h * w, dtype=_dtype_code(gcat), layout=ctx.layout, pair=1)
            _call("tlerp_cat_bwd", tlerp_bwd_bytes(B, Cs, h * w, gcat.element_size()), lib.smow_tlerp_pair_cat_bwd,
                  gcat.data_ptr(), ga.data_ptr(), gb.data_ptr(), B, Cd, Cs, h * w, _dtype_code(gcat),
                  ctx.layout, _stream())
        gdec = gcat[:, :Cd] if Cd > 0 else None
        return gdec, ga, gb


def tlerp_cat(dec, skip):
    """cat([dec, interpolate(skip, (4,h,w), trilinear, align_corners=True)], dim=1) in one launch."""
    if skip.shape[0] == 0:
        _require_cuda(dec, skip)
        cd = 0 if dec is None else dec.shape[1]
        return skip.new_zeros((0, cd + skip.shape[1], 4) + tuple(skip.shape[3:]))
    return _TLerpCat.apply(dec, skip)


def tlerp(skip):
    """Temporal 2 -> 4 upsample alone: (B,C,2,h,w) -> (B,C,4,h,w)."""
    return tlerp_cat(None, skip)


def tlerp_pair_cat(dec, x_t1, x_t2):
    """tlerp_cat on two un-stacked (B,C,h,w) frames (dec may be None)."""
    if x_t1.shape[0] == 0:
        return tlerp_cat(dec, torch.stack((x_t1, x_t2), 2))
    return _TLerpPairCat.apply(dec, x_t1, x_t2)


# ----------------------------------------------------------------------------- A3+A4 with the block's LeakyReLU folded in
def act_cat_bytes(B, Cd, hw, s):
    """The LeakyReLU pass of the decoder block (reference models/SMOW_Net.py:137), now done by the concat launch:
    z read once, the decoder half of the concat written once."""
    return B * 8 * Cd * hw * s


def act_cat_bwd_bytes(B, Cd, hw, s):
    """LeakyReLU backward of the decoder half: its gradient slice and z read once, d z written once."""
    return B * 12 * Cd * hw * s


class _ActTLerpCat(torch.autograd.Function):
    """cat([leaky_relu(z, slope), tlerp(T1, T2)], 1) in one launch; frames given as two (B,Cs,h,w) tensors or as the two
    time slices of a stacked channels_last_3d (B,Cs,2,h,w) tensor (`stacked`)."""

    @staticmethod
    def forward(ctx, z, a, b, slope, stacked):
        B, Cd, _, h, w = z.shape
        Cs = a.shape[1]
        z = z.contiguous(memory_format=torch.channels_last_3d)
        cat = torch.empty((B, Cd + Cs, 4, h, w), dtype=z.dtype, device=z.device, memory_format=torch.channels_last_3d)
        pair_stride = (2 if stacked else 1) * Cs * h * w
        lib = _lib.load()
        with torch.cuda.device_of(z):
            _meta(B=B, Cd=Cd, Cs=Cs, hw=h * w, dtype=_dtype_code(z), layout=_lib.NDHWC, pair=0, act=1)
            _call("tlerp_cat_fwd", tlerp_fwd_bytes(B, Cd, Cs, h * w, z.element_size()) + act_cat_bytes(B, Cd, h * w, z.element_size()),
                  lib.smow_act_tlerp_cat_fwd,
                  z.data_ptr(), a.data_ptr(), b.data_ptr(), cat.data_ptr(), B, Cd, Cs, h * w, pair_stride, float(slope),
                  _dtype_code(z), _stream())
        ctx.save_for_backward(z)
        ctx.cfg = (B, Cd, Cs, h, w, float(slope), stacked)
        return cat

    @staticmethod
    def backward(ctx, gcat):
        (z,) = ctx.saved_tensors
        B, Cd, Cs, h, w, slope, stacked = ctx.cfg
        gcat = gcat.contiguous(memory_format=torch.channels_last_3d)
        gz = torch.empty_like(z, memory_format=torch.channels_last_3d)
        if stacked:
            gs = torch.empty((B, Cs, 2, h, w), dtype=z.dtype, device=z.device, memory_format=torch.channels_last_3d)
            ga, gb = gs[:, :, 0], gs[:, :, 1]
        else:
            ga = torch.empty((B, Cs, h, w), dtype=z.dtype, device=z.device, memory_format=torch.channels_last)
            gb = torch.empty_like(ga, memory_format=torch.channels_last)
        pair_stride = (2 if stacked else 1) * Cs * h * w
        lib = _lib.load()
        with torch.cuda.device_of(z):
            _meta(B=B, Cd=Cd, Cs=Cs, hw=h * w, dtype=_dtype_code(z), layout=_lib.NDHWC, pair=0, act=1)
            _call("tlerp_cat_bwd", tlerp_bwd_bytes(B, Cs, h * w, z.element_size()) + act_cat_bwd_bytes(B, Cd, h * w, z.element_size()),
                  lib.smow_act_tlerp_cat_bwd,
                  gcat.data_ptr(), z.data_ptr(), gz.data_ptr(), ga.data_ptr(), gb.data_ptr(), B, Cd, Cs, h * w, pair_stride,
                  slope, _dtype_code(z), _stream())
        return gz, ga, gb, None, None


def _act_cat_ok(z, a, b):
    v = 16 // z.element_size()
    return (z.is_cuda and z.dim() == 5 and z.shape[2] == 4 and z.shape[0] > 0 and z.dtype in _DTYPES and a.dtype == z.dtype
            and b.dtype == z.dtype and z.shape[1] % v == 0 and a.shape[1] % v == 0 and tuple(a.shape) == tuple(b.shape)
            and a.shape[0] == z.shape[0] and tuple(a.shape[-2:]) == tuple(z.shape[-2:]))


def act_tlerp_pair_cat(z, x_t1, x_t2, slope=0.2):
    """``cat([leaky_relu(z, slope), interpolate(stack(x_t1, x_t2), (4,h,w))], 1)`` in ONE launch: z is the decoder block's
    BatchNorm output (B,Cd,4,h,w); the block's LeakyReLU (reference models/SMOW_Net.py:137) is applied while the decoder half
    is written into the concat buffer, so neither a stand-alone activation pass nor a copy of the decoder half remains."""
    _require_cuda(z, x_t1, x_t2)
    if not _act_cat_ok(z, x_t1, x_t2) or x_t1.dim() != 4:
        raise RuntimeError("act_tlerp_pair_cat: z (B,Cd,4,h,w) and matching (B,Cs,h,w) frames with channel counts that are "
                           "multiples of the 16-byte vector, got %s / %s" % (tuple(z.shape), tuple(x_t1.shape)))
    a = x_t1.contiguous(memory_format=torch.channels_last)
    b = x_t2.contiguous(memory_format=torch.channels_last)
    return _ActTLerpCat.apply(z, a, b, slope, False)


def act_tlerp_cat(z, skip, slope=0.2):
    """Same with the two frames stacked: skip (B,Cs,2,h,w)."""
    _require_cuda(z, skip)
    if skip.dim() != 5 or skip.shape[2] != 2 or not _act_cat_ok(z, skip[:, :, 0], skip[:, :, 1]):
        raise RuntimeError("act_tlerp_cat: z (B,Cd,4,h,w) and skip (B,Cs,2,h,w) with channel counts that are multiples of the "
                           "16-byte vector, got %s / %s" % (tuple(z.shape), tuple(skip.shape)))
    skip = skip.contiguous(memory_format=torch.channels_last_3d)
    return _ActTLerpCat.apply(z, skip[:, :, 0], skip[:, :, 1], slope, True)


def act_cat_supported(z, cs):
    v = 16 // z.element_size() if z.dtype in _DTYPES else 0
    return bool(v) and z.is_cuda and z.dim() == 5 and z.shape[2] == 4 and z.shape[0] > 0 and z.shape[1] % v == 0 and cs % v == 0


# ----------------------------------------------------------------------------- N2: semantic tokenizer
def tokenizer_fwd_bytes(B, C, hw, s=4):
    """Algorithmic bytes: the stack is read once (4 frames), tokens written."""
    return B * 4 * (C * hw * s + 8 * C * 4)


def tokenizer_bwd_bytes(B, C, hw, s=4):
    """The stack is read once and its gradient written once."""
    return B * 4 * (2 * C * hw * s + 2 * 8 * C * 4)


class _SemanticTokens(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        B, C, T, H, W = x.shape
        x = x.contiguous(memory_format=torch.channels_last_3d)
        wa = weight.reshape(weight.shape[0], C).contiguous().float()
        ba = bias.contiguous().float()
        tokens = torch.empty((B, T, 8, C), dtype=torch.float32, device=x.device)
        stats = torch.empty((B, T, 16), dtype=torch.float32, device=x.device)
        lib = _lib.load()
        n = int(lib.smow_tokenizer_workspace_bytes(B, C, H * W))
        ws = torch.empty(n, dtype=torch.uint8, device=x.device)
        with torch.cuda.device_of(x):
            _meta(B=B, C=C, hw=H * W)
            _call("tokenizer_fwd", tokenizer_fwd_bytes(B, C, H * W), lib.smow_tokenizer_fwd,
                  x.data_ptr(), wa.data_ptr(), ba.data_ptr(), tokens.data_ptr(), stats.data_ptr(), B, C, H * W,
                  _lib.F32, _lib.NDHWC, ws.data_ptr(), n, _stream())
        ctx.save_for_backward(x, wa, ba, tokens, stats)
        ctx.wshape = tuple(weight.shape)
        return tokens

    @staticmethod
    def backward(ctx, gtokens):
        x, wa, ba, tokens, stats = ctx.saved_tensors
        B, C, T, H, W = x.shape
        gtokens = gtokens.contiguous().float()
        gx = torch.empty_like(x, memory_format=torch.channels_last_3d)
        gwa, gba = torch.empty_like(wa), torch.empty_like(ba)
        lib = _lib.load()
        n = int(lib.smow_tokenizer_workspace_bytes(B, C, H * W))
        ws = torch.empty(n, dtype=torch.uint8, device=x.device)
        with torch.cuda.device_of(x):
            _meta(B=B, C=C, hw=H * W)
            _call("tokenizer_bwd", tokenizer_bwd_bytes(B, C, H * W), lib.smow_tokenizer_bwd,
                  gtokens.data_ptr(), x.data_ptr(), wa.data_ptr(), ba.data_ptr(), tokens.data_ptr(), stats.data_ptr(),
                  gx.data_ptr(), gwa.data_ptr(), gba.data_ptr(), B, C, H * W, _lib.F32, _lib.NDHWC,
                  ws.data_ptr(), n, _stream())
        return gx, gwa.view(ctx.wshape), gba


def semantic_tokens(x, weight, bias):
    """Spatial-attention pooling of every frame of the (B,C,4,H,W) stack to 8 tokens:
    ``einsum('bln,bcn->blc', softmax(conv1x1(x_k)), x_k)`` for k = 0..3 (reference models/SMOW_Net.py:176-187) in one
    pass -> (B, 4, 8, C).  ``weight`` / ``bias`` are ``conv_a``'s (8,C,1,1) and (8,)."""
    _require_cuda(x, weight, bias)
    if x.dim() != 5 or x.shape[2] != 4:
        raise RuntimeError("semantic_tokens: x must be the 4-frame stack (B,C,4,H,W), got %s" % (tuple(x.shape),))
    if weight.shape[0] != 8 or weight.shape[1] != x.shape[1] or x.dtype != torch.float32:
        raise RuntimeError("semantic_tokens: built for token_len = 8 and fp32 stacks")
    if x.shape[0] == 0:
        return x.new_zeros((0, 4, 8, x.shape[1])) + 0 * (x.sum() + weight.sum() + bias.sum())
    return _SemanticTokens.apply(x, weight, bias)


# ----------------------------------------------------------------------------- A1 + N2 fused: warp -> tokens
def warp_tokens_fwd_bytes(B, C, H, W, s=4):
    """Algorithmic bytes of the fused pass: each input frame feeds two stack frames (itself and its warp), the stack itself
    never exists: 2 x (2*C*HW*s) read + the flow + the tokens."""
    return B * (4 * C * H * W * s + 16 * H * W + 4 * 8 * C * 4)


def warp_tokens_bwd_bytes(B, C, H, W, s=4):
    """Backward of the pooling half: the rows are re-staged from x and the flow, d(stack) is written once."""
    return B * (4 * C * H * W * s + 16 * H * W + 4 * C * H * W * s + 2 * 4 * 8 * C * 4)


class _WarpTokens(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, flow, weight, bias):
        B, C, _, H, W = x.shape
        x = x.contiguous(memory_format=torch.channels_last_3d)
        ctx.flow_dtype = flow.dtype
        flow = flow.float().contiguous()
        wa = weight.reshape(weight.shape[0], C).contiguous().float()
        ba = bias.contiguous().float()
        tokens = torch.empty((B, 4, 8, C), dtype=torch.float32, device=x.device)
        stats = torch.empty((B, 4, 16), dtype=torch.float32, device=x.device)
        xs, ys = base_grid(W, x.device), base_grid(H, x.device)
        lib = _lib.load()
        n = int(lib.smow_tokenizer_workspace_bytes(B, C, H * W))
        ws = torch.empty(n, dtype=torch.uint8, device=x.device)
        with torch.cuda.device_of(x):
            _meta(B=B, C=C, H=H, W=W)
            _call("warp_tokens_fwd", warp_tokens_fwd_bytes(B, C, H, W), lib.smow_warp_tokenizer_fwd,
                  x.data_ptr(), flow.data_ptr(), xs.data_ptr(), ys.data_ptr(), wa.data_ptr(), ba.data_ptr(),
                  tokens.data_ptr(), stats.data_ptr(), B, C, H, W, _lib.F32, _lib.NDHWC, ws.data_ptr(), n, _stream())
        ctx.save_for_backward(x, flow, wa, ba, tokens, stats)
        ctx.wshape = tuple(weight.shape)
        return tokens

    @staticmethod
    def backward(ctx, gtokens):
        x, flow, wa, ba, tokens, stats = ctx.saved_tensors
        B, C, _, H, W = x.shape
        gtokens = gtokens.contiguous().float()
        gstack = _empty((B, C, 4, H, W), x, _lib.NDHWC)
        gwa, gba = torch.empty_like(wa), torch.empty_like(ba)
        gx = _empty(tuple(x.shape), x, _lib.NDHWC)
        gflow = torch.empty_like(flow)
        xs, ys = base_grid(W, x.device), base_grid(H, x.device)
        lib = _lib.load()
        n = int(lib.smow_tokenizer_workspace_bytes(B, C, H * W))
        ws = torch.empty(n, dtype=torch.uint8, device=x.device)
        fws, fws_bytes = _bwd_workspace(lib, x, _lib.NDHWC, B, H, W)
        with torch.cuda.device_of(x):
            _meta(B=B, C=C, H=H, W=W)
            _call("warp_tokens_bwd", warp_tokens_bwd_bytes(B, C, H, W), lib.smow_warp_tokenizer_bwd,
                  gtokens.data_ptr(), x.data_ptr(), flow.data_ptr(), xs.data_ptr(), ys.data_ptr(), wa.data_ptr(),
                  ba.data_ptr(), tokens.data_ptr(), stats.data_ptr(), gstack.data_ptr(), gwa.data_ptr(), gba.data_ptr(),
                  B, C, H, W, _lib.F32, _lib.NDHWC, ws.data_ptr(), n, _stream())
            _meta(B=B, C=C, H=H, W=W, dtype=_lib.F32, layout=_lib.NDHWC, pair=0)
            _call("warp_stack_bwd", warp_bwd_bytes(B, C, H, W, 4), lib.smow_warp_stack_bwd,
                  gstack.data_ptr(), x.data_ptr(), flow.data_ptr(), xs.data_ptr(), ys.data_ptr(),
                  gx.data_ptr(), gflow.data_ptr(), B, C, H, W, _lib.F32, _lib.NDHWC,
                  fws.data_ptr() if fws is not None else None, fws_bytes, _stream())
        return gx, gflow.to(ctx.flow_dtype), gwa.view(ctx.wshape), gba


def warp_tokens_supported(x, weight):
    """The fused form exists for what the two models run: fp32 (B,C,2,H,W) with C = 16 / 32, token_len = 8."""
    return bool(x.is_cuda and x.dtype == torch.float32 and x.dim() == 5 and x.shape[2] == 2 and x.shape[0] > 0
                and weight.shape[0] == 8 and _lib.load().smow_warp_tokenizer_supported(int(x.shape[1])))


def warp_tokens(x, flow, weight, bias):
    """``semantic_tokens(flow_warp(x, flow), weight, bias)`` in one pass over ``x``: the (B,C,4,H,W) stack that
    ``OFW.flow_warp`` returns (reference models/SMOW_Net.py:612-638) has a single consumer, the tokenizer of
    ``Transformer_Encoder`` (:176-187), so its rows are produced in shared memory and pooled there -> (B, 4, 8, C).
    Backward: d(stack) is produced by the same re-staged pass and handed to the warp backward."""
    _require_cuda(x, flow, weight, bias)
    if x.dim() != 5 or x.shape[2] != 2:
        raise RuntimeError("warp_tokens: input must be (B,C,2,H,W), got %s" % (tuple(x.shape),))
    B, C, _, H, W = x.shape
    if tuple(flow.shape) != (B, 2, 2, H, W):
        raise RuntimeError("warp_tokens: flow must be (B,2,2,H,W)=%s, got %s" % ((B, 2, 2, H, W), tuple(flow.shape)))
    if tuple(weight.shape[:2]) != (8, C) or bias.numel() != 8:
        raise RuntimeError("warp_tokens: conv_a must be (8,%d,1,1) / (8,), got %s / %s" % (C, tuple(weight.shape), tuple(bias.shape)))
    if not warp_tokens_supported(x, weight):
        raise RuntimeError("warp_tokens: built for fp32 stacks with C = 16 / 32 (and tok_variant != 0)")
    return _WarpTokens.apply(x, flow, weight, bias)


# ----------------------------------------------------------------------------- N4: cyclic temporal frame mix
def frame_mix_bytes(B, C, hw, s=4):
    """Algorithmic bytes of one apply pass: the 4-frame tensor read once and written once."""
    return 2 * B * 4 * C * hw * s


class _FrameMix(torch.autograd.Function):
    """y[:, :, j] = x[:, :, j] @ W_shared + x[:, :, (j+1) % 4] @ W_own[(j+1) % 4]  (matrices: rows = input channels)."""

    @staticmethod
    def forward(ctx, x, w_shared, w_own):
        B, C, T, H, W = x.shape
        x = x.contiguous(memory_format=torch.channels_last_3d)
        m0, m1 = w_shared.contiguous().float(), w_own.contiguous().float()
        y = torch.empty_like(x, memory_format=torch.channels_last_3d)
        lib = _lib.load()
        with torch.cuda.device_of(x):
            _meta(B=B, C=C, T=4, hw=H * W, tc=0)
            _call("frame_mix_fwd", frame_mix_bytes(B, C, H * W), lib.smow_frame_mix_apply,
                  x.data_ptr(), m0.data_ptr(), m1.data_ptr(), y.data_ptr(), B, C, H * W, 1, 1, _stream())
        ctx.save_for_backward(x, m0, m1)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, m0, m1 = ctx.saved_tensors
        B, C, T, H, W = x.shape
        gy = gy.contiguous(memory_format=torch.channels_last_3d)
        gx = torch.empty_like(x, memory_format=torch.channels_last_3d)
        gw = torch.empty((5, C, C), dtype=torch.float32, device=x.device)
        m0t, m1t = m0.t().contiguous(), m1.transpose(1, 2).contiguous()
        lib = _lib.load()
        n = int(lib.smow_frame_mix_wgrad_workspace_bytes(B, C, H * W))
        ws = torch.empty(n, dtype=torch.uint8, device=x.device)
        with torch.cuda.device_of(x):
            _meta(B=B, C=C, T=4, hw=H * W, tc=0)
            _call("frame_mix_bwd", frame_mix_bytes(B, C, H * W), lib.smow_frame_mix_apply,
                  gy.data_ptr(), m0t.data_ptr(), m1t.data_ptr(), gx.data_ptr(), B, C, H * W, 3, 0, _stream())
            _meta(B=B, C=C, T=4, hw=H * W, tc=0)
            _call("frame_mix_wgrad", frame_mix_wgrad_bytes(B, C, 4, H * W), lib.smow_frame_mix_wgrad,
                  x.data_ptr(), gy.data_ptr(), gw.data_ptr(), B, C, H * W, ws.data_ptr(), n, _stream())
        return gx, gw[0], gw[1:]


def frame_mix_supported(x, c_out):
    """True when the hand-written kernels take this tensor: CUDA fp32 (B,C,4,H,W) with C_in = C_out in {16,28,32,64}."""
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 5 and x.shape[2] == 4 and x.shape[1] == c_out
            and x.shape[0] > 0 and bool(_lib.load().smow_frame_mix_supported(int(c_out))))


def frame_mix(x, w_shared, w_own):
    """Cyclic temporal frame mix of the decoder blocks (reference models/SMOW_Net.py:121-139) in one pass:
    x (B,C,4,H,W), w_shared (C,C), w_own (4,C,C) with rows = input channels -> (B,C,4,H,W), channels_last_3d."""
    _require_cuda(x, w_shared, w_own)
    if w_shared.dim() != 2 or not frame_mix_supported(x, w_shared.shape[1]):
        raise RuntimeError("frame_mix: built for fp32 (B,C,4,H,W) stacks with C_in = C_out in {16, 28, 32, 64}")
    C = x.shape[1]
    if tuple(w_shared.shape) != (C, C) or tuple(w_own.shape) != (4, C, C):
        raise RuntimeError("frame_mix: w_shared must be (C,C)=%s and w_own (4,C,C), got %s and %s"
                           % ((C, C), tuple(w_shared.shape), tuple(w_own.shape)))
    if w_shared.dtype != torch.float32 or w_own.dtype != torch.float32:
        raise RuntimeError("frame_mix: fp32 matrices only")
    return _FrameMix.apply(x, w_shared, w_own)


# ----------------------------------------------------------------------------- N4 on tensor cores (tcgen05 / TMEM / TMA)
def frame_mix_apply_bytes(B, C, T, hw, s=4):
    """Algorithmic bytes of one apply pass: the T-frame tensor read once and written once."""
    return 2 * B * T * C * hw * s


def frame_mix_wgrad_bytes(B, C, T, hw, s=4):
    """Weight gradients: x and gy are read once; the (1+T) C x C results are negligible."""
    return 2 * B * T * C * hw * s + (1 + T) * C * C * 4


def frame_mix_tc_supported(x, c_out, T):
    """True when the tcgen05 kernels take this tensor: CUDA fp32 (B,C,T,H,W) channels_last_3d, C_in = C_out covered."""
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 5 and x.shape[2] == T and x.shape[1] == c_out
            and x.shape[0] > 0 and bool(_lib.load().smow_frame_mix_tc_supported(int(c_out), int(T))))


class _FrameMixTC(torch.autograd.Function):
    """out[:, :, f] = W_0 x[:, :, f] + W_{1+(f+own_off)%T} x[:, :, (f+shift)%T] (+ bias[f]) as a TMA-fed tcgen05 GEMM.

    `pack` holds the 1+T matrices either as [m][out][in] (`nk=True`: Conv3d 1x1x1 weights as stored) or as [m][in][out]
    (`nk=False`: ConvTranspose3d weights as stored); its gradient comes back in the same orientation."""

    @staticmethod
    def forward(ctx, x, pack, bias, T, shift, own_off, nk, want_stats=False):
        B, C, _, H, W = x.shape
        x = x.contiguous(memory_format=torch.channels_last_3d)
        pack = pack.contiguous()
        w_nk = pack if nk else pack.transpose(1, 2).contiguous()          # [m][out][in] feeds the forward GEMM
        y = torch.empty_like(x, memory_format=torch.channels_last_3d)
        b = None if bias is None else bias.contiguous().float()
        lib = _lib.load()
        ctx.save_for_backward(x, pack)
        ctx.cfg = (T, shift, own_off, nk, bias is not None)
        with torch.cuda.device_of(x):
            _meta(B=B, C=C, T=T, hw=H * W, tc=1)
            if want_stats:      # BatchNorm partial sums (sum y, sum y^2 per channel and CTA) from the kernel's epilogue
                parts = torch.empty((int(lib.smow_frame_mix_stats_parts(B, C, T, H * W)), 2, C), dtype=torch.float32, device=x.device)
                _call("frame_mix_fwd", frame_mix_apply_bytes(B, C, T, H * W), lib.smow_frame_mix_apply_tc_stats,
                      x.data_ptr(), w_nk.data_ptr(), None if b is None else b.data_ptr(), y.data_ptr(), parts.data_ptr(),
                      B, C, T, H * W, C, shift, own_off, _stream())
                ctx.mark_non_differentiable(parts)
                return y, parts
            _call("frame_mix_fwd", frame_mix_apply_bytes(B, C, T, H * W), lib.smow_frame_mix_apply_tc,
                  x.data_ptr(), w_nk.data_ptr(), None if b is None else b.data_ptr(), y.data_ptr(), B, C, T, H * W, C,
                  shift, own_off, _stream())
        return y

    @staticmethod
    def backward(ctx, gy, *_unused):
        x, pack = ctx.saved_tensors
        T, shift, own_off, nk, has_bias = ctx.cfg
        B, C, _, H, W = x.shape
        gy = gy.contiguous(memory_format=torch.channels_last_3d)
        w_kn = pack.transpose(1, 2).contiguous() if nk else pack           # [m][in][out] feeds the d(input) GEMM
        gx = torch.empty_like(x, memory_format=torch.channels_last_3d)
        gpack = torch.empty_like(pack)
        lib = _lib.load()
        n = int(lib.smow_frame_mix_wgrad_tc_workspace_bytes(B, C, T, H * W))
        ws = torch.empty(max(n, 16), dtype=torch.uint8, device=x.device)
        with torch.cuda.device_of(x):
            # gX_k = W_0^T gY_k + W_{1+g}^T gY_{k-shift}: the same kernel on the transposed matrices
            _meta(B=B, C=C, T=T, hw=H * W, tc=1)
            _call("frame_mix_bwd", frame_mix_apply_bytes(B, C, T, H * W), lib.smow_frame_mix_apply_tc,
                  gy.data_ptr(), w_kn.data_ptr(), None, gx.data_ptr(), B, C, T, H * W, C,
                  (T - shift) % T, (own_off - shift) % T, _stream())
            if nk:      # gradient wanted as [m][out][in]: swap the operand roles (see include/smow_b200.h)
                _meta(B=B, C=C, T=T, hw=H * W, tc=1)
                _call("frame_mix_wgrad", frame_mix_wgrad_bytes(B, C, T, H * W), lib.smow_frame_mix_wgrad_tc,
                      gy.data_ptr(), x.data_ptr(), gpack.data_ptr(), B, C, T, H * W, (T - shift) % T, (own_off - shift) % T,
                      ws.data_ptr(), n, _stream())
            else:
                _meta(B=B, C=C, T=T, hw=H * W, tc=1)
                _call("frame_mix_wgrad", frame_mix_wgrad_bytes(B, C, T, H * W), lib.smow_frame_mix_wgrad_tc,
                      x.data_ptr(), gy.data_ptr(), gpack.data_ptr(), B, C, T, H * W, shift, own_off,
                      ws.data_ptr(), n, _stream())
        gbias = gy.sum(dim=(0, 3, 4)).t() if has_bias else None             # (T, C)
        return gx, gpack, gbias, None, None, None, None, None


def frame_mix_tc(x, pack, bias=None, T=4, shift=1, own_off=1, nk=True):
    """Temporal frame mix on the 5th-generation tensor cores (TF32 products, fp32 accumulation).  x (B,C,T,H,W);
    pack (1+T,C,C); bias (T,C) or None -> (B,C,T,H,W) channels_last_3d.  Decoder blocks: T=4, shift=1, own_off=1
    (reference models/SMOW_Net.py:121-139); encoder temporal exchange: T=2, shift=1, own_off=0 (:460-473)."""
    _require_cuda(x, pack, bias)
    C = x.shape[1] if x.dim() == 5 else -1
    if not frame_mix_tc_supported(x, C, T):
        raise RuntimeError("frame_mix_tc: needs an fp32 CUDA (B,C,%d,H,W) stack with a covered channel count, got %s"
                           % (T, tuple(x.shape)))
    if tuple(pack.shape) != (1 + T, C, C) or pack.dtype != torch.float32:
        raise RuntimeError("frame_mix_tc: pack must be fp32 (1+T,C,C)=%s, got %s %s" % ((1 + T, C, C), pack.dtype, tuple(pack.shape)))
    if bias is not None and (tuple(bias.shape) != (T, C) or bias.dtype != torch.float32):
        raise RuntimeError("frame_mix_tc: bias must be fp32 (T,C)")
    return _FrameMixTC.apply(x, pack, bias, T, shift, own_off, nk, False)


def frame_mix_tc_stats_supported(x, c_out, T):
    """The statistics epilogue lives in the persistent kernel (C <= 64)."""
    return frame_mix_tc_supported(x, c_out, T) and int(_lib.load().smow_frame_mix_stats_parts(
        int(x.shape[0]), int(c_out), int(T), int(x.shape[3] * x.shape[4]))) > 0


def frame_mix_tc_stats(x, pack, bias=None, T=4, shift=1, own_off=1, nk=True):
    """frame_mix_tc that also returns the BatchNorm partial sums of its output: (y, parts[(ctas, 2, C)])."""
    _require_cuda(x, pack, bias)
    C = x.shape[1] if x.dim() == 5 else -1
    if not frame_mix_tc_stats_supported(x, C, T) or tuple(pack.shape) != (1 + T, C, C) or pack.dtype != torch.float32:
        raise RuntimeError("frame_mix_tc_stats: needs an fp32 CUDA (B,C,%d,H,W) stack with C <= 64 and a (1+T,C,C) pack" % T)
    return _FrameMixTC.apply(x, pack, bias, T, shift, own_off, nk, True)


# ----------------------------------------------------------------------------- BatchNorm + LeakyReLU + lerp + concat
def bn_act_fwd_bytes(B, Cd, hw):
    """y read once, the activated decoder half written once."""
    return B * 8 * Cd * hw * 4


def bn_act_bwd_bytes(B, Cd, hw):
    """reduction pass: gradient slice + y read; apply pass: both read again, d y written."""
    return B * 20 * Cd * hw * 4


class _BnActTLerpCat(torch.autograd.Function):
    """cat([leaky_relu(batch_norm(y)), tlerp(T1, T2)], 1) with training-mode batch statistics taken from `parts` (the
    frame-mix epilogue's partial sums): one tiny finalize launch + ONE pass forward; reduce + finalize + ONE pass backward."""

    @staticmethod
    def forward(ctx, y, parts, gamma, beta, running_mean, running_var, momentum, eps, a, b, slope, stacked):
        B, Cd, _, h, w = y.shape
        Cs = 0 if a is None else a.shape[1]
        hw = h * w
        y = y.contiguous(memory_format=torch.channels_last_3d)
        bn = torch.empty((6, Cd), dtype=torch.float32, device=y.device)
        cat = torch.empty((B, Cd + Cs, 4, h, w), dtype=torch.float32, device=y.device, memory_format=torch.channels_last_3d)
        pair_stride = (2 if stacked else 1) * Cs * hw
        lib = _lib.load()
        with torch.cuda.device_of(y):
            _lib.check(lib.smow_bn_finalize(parts.data_ptr(), parts.shape[0], Cd, B * 4 * hw, gamma.data_ptr(), beta.data_ptr(),
                                            None if running_mean is None else running_mean.data_ptr(),
                                            None if running_var is None else running_var.data_ptr(), float(momentum), float(eps),
                                            bn.data_ptr(), _stream()), "bn_finalize")
            _meta(B=B, Cd=Cd, Cs=Cs, hw=hw, dtype=_lib.F32, layout=_lib.NDHWC, pair=0, act=2)
            _call("tlerp_cat_fwd", tlerp_fwd_bytes(B, Cd, Cs, hw, 4) + bn_act_fwd_bytes(B, Cd, hw), lib.smow_bn_act_tlerp_cat_fwd,
                  y.data_ptr(), bn.data_ptr(), None if a is None else a.data_ptr(), None if b is None else b.data_ptr(),
                  cat.data_ptr(), B, Cd, Cs, hw, pair_stride, float(slope), _stream())
        ctx.save_for_backward(y, bn)
        ctx.cfg = (B, Cd, Cs, h, w, float(slope), stacked)
        return cat

    @staticmethod
    def backward(ctx, gcat):
        y, bn = ctx.saved_tensors
        B, Cd, Cs, h, w, slope, stacked = ctx.cfg
        hw = h * w
        gcat = gcat.contiguous(memory_format=torch.channels_last_3d)
        bn = bn.clone()                                              # rows 4, 5 are written by the reduction
        gy = torch.empty_like(y, memory_format=torch.channels_last_3d)
        dgamma, dbeta = torch.empty(Cd, device=y.device), torch.empty(Cd, device=y.device)
        ga = gb = None
        if Cs:
            if stacked:
                gs = torch.empty((B, Cs, 2, h, w), dtype=torch.float32, device=y.device, memory_format=torch.channels_last_3d)
                ga, gb = gs[:, :, 0], gs[:, :, 1]
            else:
                ga = torch.empty((B, Cs, h, w), dtype=torch.float32, device=y.device, memory_format=torch.channels_last)
                gb = torch.empty_like(ga, memory_format=torch.channels_last)
        pair_stride = (2 if stacked else 1) * Cs * hw
        lib = _lib.load()
        n = int(lib.smow_bn_act_bwd_workspace_bytes(B, Cd, hw))
        ws = torch.empty(max(n, 16), dtype=torch.uint8, device=y.device)
        with torch.cuda.device_of(y):
            _meta(B=B, Cd=Cd, Cs=Cs, hw=hw, dtype=_lib.F32, layout=_lib.NDHWC, pair=0, act=2)
            _call("tlerp_cat_bwd", tlerp_bwd_bytes(B, Cs, hw, 4) + bn_act_bwd_bytes(B, Cd, hw), lib.smow_bn_act_bwd_reduce,
                  gcat.data_ptr(), y.data_ptr(), bn.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), B, Cd, Cs, hw, slope,
                  ws.data_ptr(), n, _stream())
            _lib.check(lib.smow_bn_act_tlerp_cat_bwd(gcat.data_ptr(), y.data_ptr(), bn.data_ptr(), gy.data_ptr(),
                                                     None if ga is None else ga.data_ptr(), None if gb is None else gb.data_ptr(),
                                                     B, Cd, Cs, hw, pair_stride, slope, _stream()), "bn_act_tlerp_cat_bwd")
        return gy, None, dgamma, dbeta, None, None, None, None, ga, gb, None, None


def bn_act_tlerp_cat(y, parts, bn_module, slope, skip=None, skip_pair=None):
    """Training-mode ``cat([leaky_relu(bn_module(y)), interpolate(skip, (4,h,w))], 1)`` (or just the activated block output when
    no skip is given) with the batch statistics taken from `parts` — the partial sums the tcgen05 frame-mix epilogue produced
    (``frame_mix_tc_stats``).  Updates ``running_mean`` / ``running_var`` / ``num_batches_tracked`` like nn.BatchNorm3d
    (reference models/SMOW_Net.py:136-137 + :64-94)."""
    _require_cuda(y, parts, bn_module.weight, skip, *(skip_pair or ()))
    if y.dtype != torch.float32 or y.dim() != 5 or y.shape[2] != 4 or y.shape[1] % 4:
        raise RuntimeError("bn_act_tlerp_cat: y must be an fp32 (B,Cd,4,h,w) CUDA tensor with Cd % 4 == 0")
    a = b = None
    stacked = False
    if skip is not None:
        skip = skip.contiguous(memory_format=torch.channels_last_3d)
        a, b, stacked = skip[:, :, 0], skip[:, :, 1], True
    elif skip_pair is not None:
        a = skip_pair[0].contiguous(memory_format=torch.channels_last)
        b = skip_pair[1].contiguous(memory_format=torch.channels_last)
    if a is not None and (a.dtype != torch.float32 or a.shape[1] % 4 or tuple(a.shape[-2:]) != tuple(y.shape[-2:])):
        raise RuntimeError("bn_act_tlerp_cat: skip frames must be fp32 (B,Cs,h,w) with Cs % 4 == 0 matching y")
    cat = _BnActTLerpCat.apply(y, parts, bn_module.weight, bn_module.bias, bn_module.running_mean, bn_module.running_var,
                               bn_module.momentum, bn_module.eps, a, b, slope, stacked)
    if bn_module.num_batches_tracked is not None:
        bn_module.num_batches_tracked.add_(1)
    return cat


# ----------------------------------------------------------------------------- N1: the OFW flow head
def flow_head_fwd_bytes(B, C, H, W, h, w):
    """Algorithmic bytes: x read once, the low-resolution table read once, the flow written."""
    return B * (2 * C * H * W * 4 + 9 * h * w * 16 + 16 * H * W)


def flow_head_bwd_bytes(B, C, H, W, h, w):
    """d x written once, x read once (for d W), the flow gradient read, d Z written."""
    return B * (4 * C * H * W * 4 + 16 * H * W + 9 * h * w * 16)


def flow_head_supported(x, coarse):
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 5 and x.shape[2] == 2 and x.shape[0] > 0 and coarse.dim() == 5
            and bool(_lib.load().smow_flow_head_supported(int(x.shape[1]), int(x.shape[3]), int(x.shape[4]),
                                                          int(coarse.shape[3]), int(coarse.shape[4]))))


class _FlowHead(torch.autograd.Function):
    """flow = stencil(x; W[:, :C]) + sum_tap bilerp(z[tap], p + tap) — see include/smow_b200.h (row N1)."""

    @staticmethod
    def forward(ctx, x, weight, z):
        B, C, _, H, W = x.shape
        h, w = z.shape[2], z.shape[3]
        x = x.contiguous(memory_format=torch.channels_last_3d)
        weight, z = weight.contiguous(), z.contiguous()
        flow = torch.empty((B, 2, 2, H, W), dtype=torch.float32, device=x.device)
        lib = _lib.load()
        n = int(lib.smow_flow_head_workspace_bytes(B, C, H, W))
        ws = torch.empty(n, dtype=torch.uint8, device=x.device)
        with torch.cuda.device_of(x):
            _meta(B=B, C=C, H=H, W=W, h=h, w=w)
            _call("flow_head_fwd", flow_head_fwd_bytes(B, C, H, W, h, w), lib.smow_flow_head_fwd,
                  x.data_ptr(), weight.data_ptr(), z.data_ptr(), flow.data_ptr(), B, C, H, W, h, w, ws.data_ptr(), n, _stream())
        ctx.save_for_backward(x, weight)
        ctx.zshape = tuple(z.shape)
        return flow

    @staticmethod
    def backward(ctx, gflow):
        x, weight = ctx.saved_tensors
        B, C, _, H, W = x.shape
        h, w = ctx.zshape[2], ctx.zshape[3]
        gflow = gflow.contiguous().float()
        gx = torch.empty_like(x, memory_format=torch.channels_last_3d)
        gweight = torch.zeros_like(weight)                         # the kernel fills the x half; the seg half comes through z
        gz = torch.empty(ctx.zshape, dtype=torch.float32, device=x.device)
        lib = _lib.load()
        n = int(lib.smow_flow_head_workspace_bytes(B, C, H, W))
        ws = torch.empty(n, dtype=torch.uint8, device=x.device)
        with torch.cuda.device_of(x):
            _meta(B=B, C=C, H=H, W=W, h=h, w=w)
            _call("flow_head_bwd", flow_head_bwd_bytes(B, C, H, W, h, w), lib.smow_flow_head_bwd,
                  gflow.data_ptr(), x.data_ptr(), weight.data_ptr(), gx.data_ptr(), gweight.data_ptr(), gz.data_ptr(),
                  B, C, H, W, h, w, ws.data_ptr(), n, _stream())
        return gx, gweight, gz


def flow_head(x, coarse, weight):
    """The OFW flow head in one pass (reference models/SMOW_Net.py:606-608):
    ``flow_make(cat([x, interpolate(coarse, (2,H,W), trilinear, align_corners=True)], 1))`` with neither the up-sampled
    tensor nor the concat.  x (B,C,2,H,W), coarse (B,C,2,h,w) = ``down(x)``, weight (2,2C,3,3,3) -> flow (B,2,2,H,W)."""
    _require_cuda(x, coarse, weight)
    if not flow_head_supported(x, coarse):
        raise RuntimeError("flow_head: needs an fp32 CUDA (B,C,2,H,W) stack with C in {16,32,64}, W %% 4 == 0, got %s / coarse %s"
                           % (tuple(x.shape), tuple(coarse.shape)))
    B, C = x.shape[0], x.shape[1]
    if tuple(weight.shape) != (2, 2 * C, 3, 3, 3) or tuple(coarse.shape[:3]) != (B, C, 2) or weight.dtype != torch.float32 \
            or coarse.dtype != torch.float32:
        raise RuntimeError("flow_head: weight must be fp32 (2,2C,3,3,3) and coarse fp32 (B,C,2,h,w)")
    # channel contraction of the seg half at LOW resolution (up-sampling and contraction commute): output frame t sees input
    # frame t' through the temporal tap kt = t' - t + 1 (the third tap always falls on the zero padding)
    wc = weight[:, C:]
    wsel = torch.stack((wc[:, :, 1:3], wc[:, :, 0:2]), 0)                         # (t, o, c, t', kh, kw)
    z = torch.einsum("tocskl,bcsij->bklijto", wsel, coarse)                        # (B, 3, 3, h, w, 2, 2)
    z = z.reshape(B, 9, coarse.shape[3], coarse.shape[4], 4)
    return _FlowHead.apply(x, weight, z)
