#!/usr/bin/env python
"""Semantic tokenizer (row N2): hand-written kernels vs the reference's op sequence, CUDA events."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smow_net_b200 import ops
dev, cl = "cuda:0", torch.channels_last_3d


def t(fn, n=10, inner=4):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(inner): fn()
        b.record(); b.synchronize(); ts.append(a.elapsed_time(b) / inner)
    return statistics.median(ts)


def ref_tokens(x, weight, bias):
    b, c, tt, h, w = x.shape
    out = []
    for k in range(tt):
        frame = x[:, :, k]
        attn = torch.softmax(torch.nn.functional.conv2d(frame, weight, bias).reshape(b, 8, -1), dim=-1)
        out.append(torch.einsum("bln,bcn->blc", attn, frame.reshape(b, c, -1)))
    return torch.stack(out, 1)


for B, C, H in ((16, 16, 128), (16, 32, 128), (64, 32, 128), (128, 16, 128)):
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(B, C, 4, H, H, device=dev, generator=g).contiguous(memory_format=cl).requires_grad_(True)
    w = (torch.randn(8, C, 1, 1, device=dev, generator=g) / C ** 0.5).requires_grad_(True)
    bias = torch.randn(8, device=dev, generator=g).requires_grad_(True)
    gt = torch.randn(B, 4, 8, C, device=dev, generator=g)
    res = []
    for name, fn in (("ours", ops.semantic_tokens), ("aten", ref_tokens)):
        f = t(lambda: fn(x, w, bias))
        tok = fn(x, w, bias)

        def bwd():
            x.grad = w.grad = bias.grad = None
            tok.backward(gt, retain_graph=True)
        bw = t(bwd)
        fb, bb = ops.tokenizer_fwd_bytes(B, C, H * H), ops.tokenizer_bwd_bytes(B, C, H * H)
        res.append("%s fwd %.3f ms (%.0f GB/s) bwd %.3f ms (%.0f GB/s)" % (name, f, fb / f / 1e6, bw, bb / bw / 1e6))
    print("B%d C%d H%d: " % (B, C, H) + " | ".join(res), flush=True)
