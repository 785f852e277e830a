#!/usr/bin/env python
"""Isolated hot-path kernel sweep (BASELINE.json configs[4]): warp+stack fwd/bwd and tlerp+concat
fwd/bwd, HBM-cold (working set >= --min-bytes, several times the 126 MB L2), CUDA-event timed, against the
reference's own op sequence on the same GPU (benchmarks/_aten_baseline.py: F.grid_sample / F.interpolate /
torch.cat = ATen's sm_100 kernels).

    python benchmarks/sweep_warp.py [--quick] [--out gpurun_out/sweep.jsonl]

One JSON line per (op, shape, dtype, variant): ms (median), algorithmic GB/s, fraction of the measured
HBM peak, and the speed-up over the ATen sequence.
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch  # noqa: E402

from _aten_baseline import aten_flow_warp, aten_tlerp_cat  # noqa: E402
from smow_net_b200 import _lib, ops  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


INNER = 4   # calls enqueued between one event pair, so host launch latency does not pad sub-100 us kernels


def time_fn(fn, warm, iters):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(INNER):
            fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) / INNER)
    return statistics.median(ts)


def pick_batch(bytes_per_pair, min_bytes, cap=512):
    return int(max(1, min(cap, -(-min_bytes // bytes_per_pair))))


def sweep_warp(args, emit):
    dev = "cuda:0"
    shapes = [(16, 128), (32, 128)] + [(c, h) for c in (64, 128, 256) for h in (64, 128, 256)]
    if args.quick:
        shapes = [(16, 128), (32, 128), (64, 128), (256, 64)]
    for dtype in (torch.float32, torch.bfloat16):
        s = 4 if dtype == torch.float32 else 2
        for C, H in shapes:
            W = H
            for sigma in (0.3, 8.0):
                if sigma == 8.0 and (args.quick or C not in (32, 64)):
                    continue
                per_pair = ops.warp_bwd_bytes(1, C, H, W, s)
                B = pick_batch(per_pair, args.min_bytes)
                g = torch.Generator(device=dev).manual_seed(1)
                x = torch.randn(B, C, 2, H, W, device=dev, generator=g).to(dtype)
                flow = torch.randn(B, 2, 2, H, W, device=dev, generator=g) * sigma
                gout = torch.randn(B, C, 4, H, W, device=dev, generator=g).to(dtype)
                base = {"C": C, "H": H, "W": W, "B": B, "dtype": str(dtype).split(".")[-1], "sigma": sigma}
                # reference sequence (fp32 only: its bf16 path quantises the grid)
                ref_f = ref_b = None
                if dtype == torch.float32:
                    xr, fr = x.clone().requires_grad_(True), flow.clone().requires_grad_(True)
                    with torch.no_grad():
                        ref_f = time_fn(lambda: aten_flow_warp(x, flow), args.warm, args.iters)

                    def ref_fb():
                        xr.grad = fr.grad = None
                        aten_flow_warp(xr, fr).backward(gout)
                    ref_b = time_fn(ref_fb, args.warm, args.iters) - ref_f
                for fv in args.fwd_variants:
                    _lib.set_option("warp_fwd_variant", fv)
                    with torch.no_grad():
                        ms = time_fn(lambda: ops.flow_warp(x, flow, (H, W)), args.warm, args.iters)
                    nb = ops.warp_fwd_bytes(B, C, H, W, s)
                    emit(dict(base, op="warp_stack_fwd", variant=fv, ms=ms, gbps=nb / ms / 1e6,
                              frac=nb / ms / 1e6 / peak(), ref_ms=ref_f, speedup=(ref_f / ms) if ref_f else None))
                xg, fg = x.clone().requires_grad_(True), flow.clone().requires_grad_(True)
                out = ops.flow_warp(xg, fg, (H, W))
                for bv in args.bwd_variants:
                    _lib.set_option("warp_bwd_variant", bv)

                    def bwd():
                        xg.grad = fg.grad = None
                        out.backward(gout, retain_graph=True)
                    ms = time_fn(bwd, args.warm, args.iters)
                    nb = ops.warp_bwd_bytes(B, C, H, W, s)
                    emit(dict(base, op="warp_stack_bwd", variant=bv, ms=ms, gbps=nb / ms / 1e6,
                              frac=nb / ms / 1e6 / peak(), ref_ms=ref_b, speedup=(ref_b / ms) if ref_b else None))
                if dtype == torch.float32:      # channels_last_3d tensors -> NDHWC kernels (reported as variant 9)
                    cl = torch.channels_last_3d
                    xc = x.contiguous(memory_format=cl)
                    gc = gout.contiguous(memory_format=cl)
                    with torch.no_grad():
                        ms = time_fn(lambda: ops.flow_warp(xc, flow, (H, W)), args.warm, args.iters)
                    nb = ops.warp_fwd_bytes(B, C, H, W, s)
                    emit(dict(base, op="warp_stack_fwd", variant=9, ms=ms, gbps=nb / ms / 1e6, frac=nb / ms / 1e6 / peak(),
                              ref_ms=ref_f, speedup=(ref_f / ms) if ref_f else None))
                    xcg, fcg = xc.clone(memory_format=torch.preserve_format).requires_grad_(True), flow.clone().requires_grad_(True)
                    outc = ops.flow_warp(xcg, fcg, (H, W))

                    def bwdc():
                        xcg.grad = fcg.grad = None
                        outc.backward(gc, retain_graph=True)
                    nb = ops.warp_bwd_bytes(B, C, H, W, s)
                    # variant 9 = default NDHWC backward (tile gather + far pass), 8 = vector-atomic scatter
                    for tag, opt in ((9, -1), (8, 0)):
                        _lib.set_option("warp_bwd_variant", opt)
                        ms = time_fn(bwdc, args.warm, args.iters)
                        emit(dict(base, op="warp_stack_bwd", variant=tag, ms=ms, gbps=nb / ms / 1e6,
                                  frac=nb / ms / 1e6 / peak(), ref_ms=ref_b, speedup=(ref_b / ms) if ref_b else None))
                    _lib.set_option("warp_bwd_variant", -1)
                    del xc, gc, xcg, fcg, outc
                del x, flow, gout, out, xg, fg
                torch.cuda.empty_cache()


def sweep_tokenizer(args, emit):
    """Row N2: semantic tokenizer vs the reference's per-frame conv + softmax + einsum sequence."""
    dev, cl = "cuda:0", torch.channels_last_3d
    for C, H, B in ((16, 128, 16), (16, 128, 128), (32, 128, 16), (32, 128, 64)):
        g = torch.Generator(device=dev).manual_seed(3)
        x = torch.randn(B, C, 4, H, H, device=dev, generator=g).contiguous(memory_format=cl).requires_grad_(True)
        w = (torch.randn(8, C, 1, 1, device=dev, generator=g) / C ** 0.5).requires_grad_(True)
        bias = torch.randn(8, device=dev, generator=g).requires_grad_(True)
        gt = torch.randn(B, 4, 8, C, device=dev, generator=g)

        def ref(x, w, bias):
            out = []
            for k in range(4):
                f = x[:, :, k]
                a = torch.softmax(torch.nn.functional.conv2d(f, w, bias).reshape(B, 8, -1), dim=-1)
                out.append(torch.einsum("bln,bcn->blc", a, f.reshape(B, C, -1)))
            return torch.stack(out, 1)
        base = {"C": C, "H": H, "W": H, "B": B, "dtype": "float32", "sigma": 0.0}
        res = {}
        for name, fn in (("ref", ref), ("ours", ops.semantic_tokens)):
            with torch.no_grad():
                f = time_fn(lambda: fn(x, w, bias), args.warm, args.iters)
            tok = fn(x, w, bias)

            def bwd():
                x.grad = w.grad = bias.grad = None
                tok.backward(gt, retain_graph=True)
            res[name] = (f, time_fn(bwd, args.warm, args.iters))
        for i, (op, nb) in enumerate((("tokenizer_fwd", ops.tokenizer_fwd_bytes(B, C, H * H)),
                                      ("tokenizer_bwd", ops.tokenizer_bwd_bytes(B, C, H * H)))):
            ms, rms = res["ours"][i], res["ref"][i]
            emit(dict(base, op=op, variant=9, ms=ms, gbps=nb / ms / 1e6, frac=nb / ms / 1e6 / peak(), ref_ms=rms,
                      speedup=rms / ms))
        del x, tok
        torch.cuda.empty_cache()


def sweep_tlerp(args, emit):
    dev = "cuda:0"
    # (Cd, Cs, h): the two models' five decoder levels
    shapes = [(32, 32, 128), (64, 32, 64), (64, 64, 32), (128, 128, 16), (256, 256, 8), (28, 16, 128), (32, 24, 64)]
    if args.quick:
        shapes = shapes[:2]
    for dtype in (torch.float32, torch.bfloat16):
        s = 4 if dtype == torch.float32 else 2
        for Cd, Cs, h in shapes:
            per_pair = ops.tlerp_fwd_bytes(1, Cd, Cs, h * h, s)
            B = pick_batch(per_pair, args.min_bytes, cap=8192)
            g = torch.Generator(device=dev).manual_seed(2)
            skip = torch.randn(B, Cs, 2, h, h, device=dev, generator=g).to(dtype)
            dec = torch.randn(B, Cd, 4, h, h, device=dev, generator=g).to(dtype)
            gcat = torch.randn(B, Cd + Cs, 4, h, h, device=dev, generator=g).to(dtype)
            base = {"Cd": Cd, "Cs": Cs, "h": h, "B": B, "dtype": str(dtype).split(".")[-1]}
            with torch.no_grad():
                ref_f = time_fn(lambda: aten_tlerp_cat(dec, skip), args.warm, args.iters)
                ms = time_fn(lambda: ops.tlerp_cat(dec, skip), args.warm, args.iters)
            nb = ops.tlerp_fwd_bytes(B, Cd, Cs, h * h, s)
            emit(dict(base, op="tlerp_cat_fwd", variant=0, ms=ms, gbps=nb / ms / 1e6, frac=nb / ms / 1e6 / peak(),
                      ref_ms=ref_f, speedup=ref_f / ms))
            sg = skip.clone().requires_grad_(True)
            cat = ops.tlerp_cat(dec, sg)

            def bwd():
                sg.grad = None
                cat.backward(gcat, retain_graph=True)
            ms = time_fn(bwd, args.warm, args.iters)
            sr = skip.clone().requires_grad_(True)
            catr = aten_tlerp_cat(dec, sr)

            def rbwd():
                sr.grad = None
                catr.backward(gcat, retain_graph=True)
            ref_b = time_fn(rbwd, args.warm, args.iters)
            nb = ops.tlerp_bwd_bytes(B, Cs, h * h, s)
            emit(dict(base, op="tlerp_cat_bwd", variant=0, ms=ms, gbps=nb / ms / 1e6, frac=nb / ms / 1e6 / peak(),
                      ref_ms=ref_b, speedup=ref_b / ms))
            cl = torch.channels_last_3d     # channels_last_3d tensors -> NDHWC kernels (variant 9)
            v = 16 // s
            if Cd % v == 0 and Cs % v == 0:
                skc, dc, gc = (t.contiguous(memory_format=cl) for t in (skip, dec, gcat))
                with torch.no_grad():
                    ms = time_fn(lambda: ops.tlerp_cat(dc, skc), args.warm, args.iters)
                nb = ops.tlerp_fwd_bytes(B, Cd, Cs, h * h, s)
                emit(dict(base, op="tlerp_cat_fwd", variant=9, ms=ms, gbps=nb / ms / 1e6, frac=nb / ms / 1e6 / peak(),
                          ref_ms=ref_f, speedup=ref_f / ms))
                sgc = skc.clone(memory_format=torch.preserve_format).requires_grad_(True)
                catc = ops.tlerp_cat(dc, sgc)

                def bwdc():
                    sgc.grad = None
                    catc.backward(gc, retain_graph=True)
                ms = time_fn(bwdc, args.warm, args.iters)
                nb = ops.tlerp_bwd_bytes(B, Cs, h * h, s)
                emit(dict(base, op="tlerp_cat_bwd", variant=9, ms=ms, gbps=nb / ms / 1e6, frac=nb / ms / 1e6 / peak(),
                          ref_ms=ref_b, speedup=ref_b / ms))
                del skc, dc, gc, sgc, catc
            del skip, dec, gcat, cat, catr
            torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="", choices=["", "warp", "tlerp", "tokenizer"])
    ap.add_argument("--min-bytes", type=int, default=1 << 30)
    ap.add_argument("--warm", type=int, default=5)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--bwd-variants", type=int, nargs="+", default=[0, 1, 2])
    ap.add_argument("--fwd-variants", type=int, nargs="+", default=[0, 1, 2])
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    fh = open(args.out, "w")

    def emit(rec):
        line = json.dumps(rec)
        fh.write(line + "\n")
        fh.flush()
        print(line)
    saved = {k: _lib.get_option(k) for k in ("warp_fwd_variant", "warp_bwd_variant")}
    if args.only in ("", "warp"):
        sweep_warp(args, emit)
    if args.only in ("", "tlerp"):
        sweep_tlerp(args, emit)
    if args.only in ("", "tokenizer"):
        sweep_tokenizer(args, emit)
    for k, v in saved.items():
        _lib.set_option(k, v)


if __name__ == "__main__":
    main()
