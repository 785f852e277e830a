#!/usr/bin/env python
"""Semantic tokenizer (row N2): tensor-core MMA kernels (tok_variant = -1) vs the FP32-pipe kernels (tok_variant = 0).

Accuracy of both against a float64 evaluation of the reference's op sequence, then HBM-cold graph-replayed timings
(smow_net_b200.probe) of the C-ABI calls."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from smow_net_b200 import _lib, ops, probe

dev, cl = torch.device("cuda:0"), torch.channels_last_3d


def ref_tokens(x, weight, bias):
    b, c, tt, h, w = x.shape
    out = []
    for k in range(tt):
        frame = x[:, :, k]
        attn = torch.softmax(torch.nn.functional.conv2d(frame, weight, bias).reshape(b, 8, -1), dim=-1)
        out.append(torch.einsum("bln,bcn->blc", attn, frame.reshape(b, c, -1)))
    return torch.stack(out, 1)


def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))


for B, C, H, W in ((2, 16, 7, 9), (3, 32, 25, 40), (2, 16, 64, 64), (2, 32, 128, 128)):
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(B, C, 4, H, W, device=dev, generator=g).contiguous(memory_format=cl)
    w = torch.randn(8, C, 1, 1, device=dev, generator=g) / C ** 0.5
    bias = torch.randn(8, device=dev, generator=g)
    gt = torch.randn(B, 4, 8, C, device=dev, generator=g)
    xd, wd, bd = (v.double().requires_grad_(True) for v in (x, w, bias))
    tok_ref = ref_tokens(xd, wd, bd)
    tok_ref.backward(gt.double())
    line = []
    for variant in (0, -1):
        _lib.set_option("tok_variant", variant)
        xx, ww, bb = (v.clone().requires_grad_(True) for v in (x, w, bias))
        tok = ops.semantic_tokens(xx, ww, bb)
        tok.backward(gt)
        line.append("variant %2d: tok %.1e gx %.1e gw %.1e gb %.1e" % (
            variant, rel(tok, tok_ref.detach()), rel(xx.grad, xd.grad), rel(ww.grad, wd.grad), rel(bb.grad, bd.grad)))
    print("B%d C%d %dx%d | " % (B, C, H, W) + " | ".join(line), flush=True)

for B, C in ((16, 16), (16, 32), (64, 32)):
    for name in ("tokenizer_fwd", "tokenizer_bwd"):
        out = []
        for variant in (0, -1):
            _lib.set_option("tok_variant", variant)
            r = probe.time_call(name, {"B": B, "C": C, "hw": 128 * 128}, dev)
            out.append("variant %2d %.1f us cold (%.0f GB/s) %.1f us warm" % (variant, r["cold_ms"] * 1e3, r["bytes"] / r["cold_ms"] / 1e6,
                                                                              r["warm_ms"] * 1e3))
            torch.cuda.empty_cache()
        print("%s B%d C%d: " % (name, B, C) + " | ".join(out), flush=True)
