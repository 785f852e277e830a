#!/usr/bin/env python
"""Run each hot-path kernel a few times on one HBM-cold shape (ncu target; also prints CUDA-event times).

    python benchmarks/one_kernel.py --C 32 --H 128 --B 64 --iters 3 [--dtype bf16] [--sigma 0.3]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from smow_net_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--C", type=int, default=32)
    ap.add_argument("--H", type=int, default=128)
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--sigma", type=float, default=0.3)
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--layout", default="ncdhw", choices=["ncdhw", "ndhwc"])
    a = ap.parse_args()
    dt = torch.float32 if a.dtype == "f32" else torch.bfloat16
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(a.B, a.C, 2, a.H, a.H, device=dev, generator=g).to(dt).requires_grad_(True)
    flow = (torch.randn(a.B, 2, 2, a.H, a.H, device=dev, generator=g) * a.sigma).requires_grad_(True)
    gout = torch.randn(a.B, a.C, 4, a.H, a.H, device=dev, generator=g).to(dt)
    dec = torch.randn(a.B, a.C, 4, a.H, a.H, device=dev, generator=g).to(dt)
    if a.layout == "ndhwc":
        cl = torch.channels_last_3d
        x = x.detach().contiguous(memory_format=cl).requires_grad_(True)
        gout, dec = gout.contiguous(memory_format=cl), dec.contiguous(memory_format=cl)
    for i in range(a.iters):
        with ops.kernel_timer() as kt:
            out = ops.flow_warp(x, flow, (a.H, a.H))
            out.backward(gout)
            x.grad = flow.grad = None
            cat = ops.tlerp_cat(dec, x)
            cat.backward(torch.cat([gout, gout], 1).contiguous(memory_format=torch.channels_last_3d if a.layout == 'ndhwc' else torch.contiguous_format))
            x.grad = None
            if a.layout == "ndhwc" and dt == torch.float32:
                w = (torch.randn(8, a.C, 1, 1, device=dev, generator=g) / a.C ** 0.5).requires_grad_(True)
                bias = torch.zeros(8, device=dev, requires_grad=True)
                stack = dec.detach().requires_grad_(True)
                ops.semantic_tokens(stack, w, bias).backward(torch.randn(a.B, 4, 8, a.C, device=dev, generator=g))
                if ops.frame_mix_supported(dec, a.C):      # row N4: cyclic frame mix (apply fwd, apply d-input, weight grads)
                    ws_ = (torch.randn(a.C, a.C, device=dev, generator=g) / a.C ** 0.5).requires_grad_(True)
                    wo_ = (torch.randn(4, a.C, a.C, device=dev, generator=g) / a.C ** 0.5).requires_grad_(True)
                    ops.frame_mix(dec.detach().requires_grad_(True), ws_, wo_).backward(gout)
            torch.cuda.synchronize()
        print(i, {k: "%.3f ms %.0f GB/s" % (v["ms"], v["gbps"]) for k, v in kt.summary().items()})


if __name__ == "__main__":
    main()
