"""numpy front-end of oracle/smow_oracle.c (TEST INFRASTRUCTURE ONLY)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "smow_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB_PATH)
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def linspace_table(n):
    """fp32 base-grid table exactly as the reference builds it: torch.linspace(-1, 1, n) on the CPU
    (models/SMOW_Net.py:617-618).  ATen's CPU kernel is vectorised (base + step*lane per SIMD
    group), so its low bits are not those of a scalar formula; the table is therefore taken from
    torch itself, like the product does (smow_net_b200.ops.base_grid)."""
    import torch
    return torch.linspace(-1.0, 1.0, int(n)).numpy().astype(np.float32)


def warp_stack_fwd(x, flow, xs=None, ys=None):
    x, flow = _f32(x), _f32(flow)
    B, C, T, H, W = x.shape
    assert T == 2 and flow.shape == (B, 2, 2, H, W)
    xs = _f32(linspace_table(W) if xs is None else xs)
    ys = _f32(linspace_table(H) if ys is None else ys)
    out = np.empty((B, C, 4, H, W), np.float32)
    _load().smow_oracle_warp_stack_fwd(_p(x), _p(flow), _p(xs), _p(ys), _p(out), B, C, H, W)
    return out


def warp_stack_bwd(gout, x, flow, xs=None, ys=None):
    gout, x, flow = _f32(gout), _f32(x), _f32(flow)
    B, C, T, H, W = x.shape
    xs = _f32(linspace_table(W) if xs is None else xs)
    ys = _f32(linspace_table(H) if ys is None else ys)
    gx = np.empty_like(x)
    gflow = np.empty_like(flow)
    _load().smow_oracle_warp_stack_bwd(_p(gout), _p(x), _p(flow), _p(xs), _p(ys), _p(gx), _p(gflow), B, C, H, W)
    return gx, gflow


def tlerp_cat_fwd(dec, skip):
    skip = _f32(skip)
    B, Cs, T, h, w = skip.shape
    Cd = 0
    if dec is not None:
        dec = _f32(dec)
        Cd = dec.shape[1]
    cat = np.empty((B, Cd + Cs, 4, h, w), np.float32)
    _load().smow_oracle_tlerp_cat_fwd(_p(dec), _p(skip), _p(cat), B, Cd, Cs, ctypes.c_int64(h * w))
    return cat


def tlerp_cat_bwd(gcat, Cd):
    gcat = _f32(gcat)
    B, Ct, T, h, w = gcat.shape
    Cs = Ct - Cd
    gskip = np.empty((B, Cs, 2, h, w), np.float32)
    _load().smow_oracle_tlerp_cat_bwd(_p(gcat), _p(gskip), B, Cd, Cs, ctypes.c_int64(h * w))
    return gskip


def tokenizer_fwd(x, wa, ba):
    """Row N2 (reference models/SMOW_Net.py:176-187): x (B,C,4,H,W), wa (L,C[,1,1]), ba (L,) -> tokens (B,4,L,C)."""
    x, wa, ba = _f32(x), _f32(np.asarray(wa).reshape(np.asarray(wa).shape[0], -1)), _f32(ba)
    B, C, T, H, W = x.shape
    assert T == 4 and wa.shape[1] == C and C <= 512
    L = wa.shape[0]
    tokens = np.empty((B, 4, L, C), np.float32)
    _load().smow_oracle_tokenizer_fwd(_p(x), _p(wa), _p(ba), _p(tokens), B, C, L, ctypes.c_int64(H * W))
    return tokens


def frame_mix_fwd(x, w_shared, w_own, bias=None):
    """Row N4 (reference models/SMOW_Net.py:121-139): x (B,Cin,4,H,W), w_shared (Cin,Cout), w_own (4,Cin,Cout) with
    rows = input channels, bias (4,Cout) or None -> (B,Cout,4,H,W)."""
    x, ws, wo = _f32(x), _f32(w_shared), _f32(w_own)
    B, Cin, T, H, W = x.shape
    Cout = ws.shape[1]
    assert T == 4 and ws.shape == (Cin, Cout) and wo.shape == (4, Cin, Cout)
    bias = None if bias is None else _f32(bias)
    out = np.empty((B, Cout, 4, H, W), np.float32)
    _load().smow_oracle_frame_mix_fwd(_p(x), _p(ws), _p(wo), _p(bias), _p(out), B, Cin, Cout, ctypes.c_int64(H * W))
    return out
