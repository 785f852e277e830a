#!/usr/bin/env python
"""GPU probe of the tcgen05 frame-mix kernels against a torch fp64 evaluation of the same formula (one process per case
group, so a trapped launch cannot poison the next group).

    python benchmarks/tc_probe.py apply 0      # correctness cases of smow_frame_mix_apply_tc, group 0..N
    python benchmarks/tc_probe.py time         # timing vs the SIMT kernel at the models' shapes
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from smow_net_b200 import _lib

dev = torch.device("cuda:0")
lib = _lib.load()
CL = torch.channels_last_3d


def ref_mix(x, wpack, bias, T, shift, own_off):
    """x (B,C,T,H,W); wpack (1+T, Cout, Cin); fp64."""
    xd, wd = x.double(), wpack.double()
    out = []
    for f in range(T):
        y = torch.einsum("bchw,dc->bdhw", xd[:, :, f], wd[0]) + \
            torch.einsum("bchw,dc->bdhw", xd[:, :, (f + shift) % T], wd[1 + (f + own_off) % T])
        if bias is not None:
            y = y + bias[f].double().view(1, -1, 1, 1)
        out.append(y)
    return torch.stack(out, 2)


def run_apply(B, C, T, H, W, pitch_extra, use_bias, shift, own_off, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(B, C, T, H, W, device=dev, generator=g).contiguous(memory_format=CL)
    wpack = torch.randn(1 + T, C, C, device=dev, generator=g) / C ** 0.5
    bias = torch.randn(T, C, device=dev, generator=g) if use_bias else None
    pitch = C + pitch_extra
    buf = torch.full((B, pitch, T, H, W), 7.0, device=dev).contiguous(memory_format=CL)
    rc = lib.smow_frame_mix_apply_tc(x.data_ptr(), wpack.data_ptr(), bias.data_ptr() if use_bias else None, buf.data_ptr(),
                                     B, C, T, H * W, pitch, shift, own_off, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "apply_tc")
    torch.cuda.synchronize()
    want = ref_mix(x, wpack, bias, T, shift, own_off)
    got = buf[:, :C].double()
    err = float((got - want).abs().max())
    scale = float(want.abs().max())
    untouched = bool((buf[:, C:] == 7.0).all()) if pitch_extra else True
    return err, scale, untouched


def run_wgrad(B, C, T, H, W, shift, own_off, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(B, C, T, H, W, device=dev, generator=g).contiguous(memory_format=CL)
    gy = torch.randn(B, C, T, H, W, device=dev, generator=g).contiguous(memory_format=CL)
    gw = torch.full((1 + T, C, C), 7.0, device=dev)
    n = int(lib.smow_frame_mix_wgrad_tc_workspace_bytes(B, C, T, H * W))
    ws = torch.empty(max(n, 16), dtype=torch.uint8, device=dev)
    _lib.check(lib.smow_frame_mix_wgrad_tc(x.data_ptr(), gy.data_ptr(), gw.data_ptr(), B, C, T, H * W, shift, own_off,
                                           ws.data_ptr(), n, torch.cuda.current_stream().cuda_stream), "wgrad_tc")
    torch.cuda.synchronize()
    xd, gd = x.double(), gy.double()
    want = torch.zeros(1 + T, C, C, dtype=torch.float64, device=dev)
    for f in range(T):
        want[0] += torch.einsum("bchw,bdhw->cd", xd[:, :, f], gd[:, :, f])
        want[1 + (f + own_off) % T] += torch.einsum("bchw,bdhw->cd", xd[:, :, (f + shift) % T], gd[:, :, f])
    err = float((gw.double() - want).abs().max())
    return err, float(want.abs().max())


WGRAD_GROUPS = [
    [(2, 32, 4, 16, 16, 1, 1), (1, 32, 4, 8, 8, 1, 1), (3, 32, 4, 20, 12, 1, 1), (16, 32, 4, 64, 64, 1, 1)],
    [(2, 16, 4, 16, 16, 1, 1), (2, 28, 4, 16, 16, 1, 1), (2, 64, 4, 16, 16, 1, 1), (2, 64, 2, 32, 32, 1, 0)],
    [(1, 128, 4, 16, 8, 1, 1), (1, 160, 4, 8, 8, 1, 1), (1, 256, 4, 8, 8, 1, 1), (1, 320, 4, 8, 8, 1, 1), (1, 512, 2, 8, 8, 1, 0)],
]

APPLY_GROUPS = [
    [(2, 32, 4, 16, 16, 0, False, 1, 1), (2, 32, 4, 16, 16, 24, True, 1, 1), (1, 32, 4, 8, 8, 0, True, 3, 0),
     (3, 32, 4, 20, 12, 0, False, 1, 1)],
    [(2, 16, 4, 16, 16, 0, False, 1, 1), (2, 16, 4, 16, 16, 28, True, 1, 1), (2, 28, 4, 16, 16, 0, True, 1, 1),
     (2, 28, 4, 16, 16, 16, False, 3, 0)],
    [(2, 64, 4, 16, 16, 0, True, 1, 1), (1, 128, 4, 16, 8, 128, False, 1, 1), (1, 160, 4, 8, 8, 0, True, 3, 0)],
    [(1, 256, 4, 8, 8, 0, False, 1, 1), (1, 320, 4, 8, 8, 320, True, 1, 1), (1, 512, 2, 8, 8, 0, False, 1, 0),
     (2, 64, 2, 32, 32, 0, False, 1, 0)],
]


def main():
    mode = sys.argv[1]
    if mode == "apply":
        grp = int(sys.argv[2])
        for case in APPLY_GROUPS[grp]:
            try:
                err, scale, ok = run_apply(*case)
                print("apply B%d C%d T%d %dx%d pitch+%d bias=%s shift=%d off=%d : max err %.3e (scale %.2f, rel %.2e) untouched=%s %s"
                      % (*case, err, scale, err / scale, ok, "OK" if err / scale < 3e-3 and ok else "FAIL"), flush=True)
            except Exception as e:
                print("apply %s : EXCEPTION %s" % (case, str(e).splitlines()[0][:200]), flush=True)
                break
    elif mode == "wgrad":
        grp = int(sys.argv[2])
        for case in WGRAD_GROUPS[grp]:
            try:
                err, scale = run_wgrad(*case)
                print("wgrad B%d C%d T%d %dx%d shift=%d off=%d : max err %.3e (scale %.2f, rel %.2e) %s"
                      % (*case, err, scale, err / scale, "OK" if err / scale < 3e-3 else "FAIL"), flush=True)
            except Exception as e:
                print("wgrad %s : EXCEPTION %s" % (case, str(e).splitlines()[0][:200]), flush=True)
                break
    elif mode == "wtime":
        for B, C, H in ((16, 16, 128), (16, 32, 128), (64, 32, 128), (16, 64, 64), (16, 28, 128)):
            g = torch.Generator(device=dev).manual_seed(0)
            x = torch.randn(B, C, 4, H, H, device=dev, generator=g).contiguous(memory_format=CL)
            gy = torch.randn(B, C, 4, H, H, device=dev, generator=g).contiguous(memory_format=CL)
            gw = torch.empty(5, C, C, device=dev)
            st = torch.cuda.current_stream().cuda_stream
            n1 = int(lib.smow_frame_mix_wgrad_tc_workspace_bytes(B, C, 4, H * H))
            n2 = int(lib.smow_frame_mix_wgrad_workspace_bytes(B, C, H * H))
            ws = torch.empty(max(n1, n2), dtype=torch.uint8, device=dev)

            def tc():
                _lib.check(lib.smow_frame_mix_wgrad_tc(x.data_ptr(), gy.data_ptr(), gw.data_ptr(), B, C, 4, H * H, 1, 1,
                                                       ws.data_ptr(), n1, st), "tc")

            def simt():
                _lib.check(lib.smow_frame_mix_wgrad(x.data_ptr(), gy.data_ptr(), gw.data_ptr(), B, C, H * H, ws.data_ptr(), n2, st), "simt")
            res = []
            for name, fn in (("tcgen05", tc), ("simt", simt)):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 20
                e0.record()
                for _ in range(n):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / n
                res.append("%s %.1f us %.0f GB/s" % (name, ms * 1e3, 2 * x.numel() * 4 / ms / 1e6))
            print("wgrad B%d C%d %dx%d: %s" % (B, C, H, H, " | ".join(res)), flush=True)
    elif mode == "time":
        for B, C, H in ((16, 16, 128), (16, 32, 128), (64, 32, 128), (16, 64, 64), (16, 28, 128)):
            g = torch.Generator(device=dev).manual_seed(0)
            x = torch.randn(B, C, 4, H, H, device=dev, generator=g).contiguous(memory_format=CL)
            wpack = torch.randn(5, C, C, device=dev, generator=g) / C ** 0.5
            y = torch.empty_like(x)
            st = torch.cuda.current_stream().cuda_stream

            def tc():
                _lib.check(lib.smow_frame_mix_apply_tc(x.data_ptr(), wpack.data_ptr(), None, y.data_ptr(), B, C, 4, H * H, C,
                                                       1, 1, st), "tc")

            def simt():
                _lib.check(lib.smow_frame_mix_apply(x.data_ptr(), wpack[0].data_ptr(), wpack[1:].data_ptr(), y.data_ptr(), B, C,
                                                    H * H, 1, 1, st), "simt")
            res = []
            for name, fn in (("tcgen05", tc), ("simt", simt)):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 20
                e0.record()
                for _ in range(n):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / n
                res.append("%s %.1f us %.0f GB/s" % (name, ms * 1e3, 2 * x.numel() * 4 / ms / 1e6))
            print("B%d C%d %dx%d: %s" % (B, C, H, H, " | ".join(res)), flush=True)


if __name__ == "__main__":
    main()
