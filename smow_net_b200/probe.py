"""Stand-alone timing of the hand-written kernels through the C ABI (used by bench.py's `roofline` / `sweep` blocks and by
benchmarks/): every call a training step made is rebuilt from its recorded shape (`ops.kernel_timer().calls()`), given
fresh operands, captured in a CUDA graph and replayed, so Python / launch latency is out of the picture.

Two regimes, both on the launching stream with CUDA events around graph replays:
  cold   the operands rotate over enough independent buffer sets that the footprint exceeds `footprint` bytes
         (default 1 GiB, several times the 126 MB L2): HBM-cold, what SURVEY §8(d)'s roofline asks for;
  warm   one buffer set, replayed back to back (operands L2-resident when they fit): a lower bound of the in-step time.
No oracle, no PyTorch fallback: the calls are the C ABI of libsmow_b200.so.
"""
import torch

from . import _lib, ops

CL3 = torch.channels_last_3d
CL2 = torch.channels_last


def _t(shape, dev, dtype, layout, gen, scale=1.0):
    t = torch.randn(shape, device=dev, generator=gen) * scale
    t = t.to(dtype)
    if layout == _lib.NDHWC:
        t = t.contiguous(memory_format=CL3 if len(shape) == 5 else CL2)
    return t


def _dt(code):
    return torch.float32 if code == _lib.F32 else torch.bfloat16


def build(name, m, dev, gen, sigma=0.3):
    """-> (callable that enqueues the C-ABI call on the current stream, algorithmic bytes, operand bytes, keepalive)."""
    lib = _lib.load()
    st = lambda: torch.cuda.current_stream().cuda_stream    # noqa: E731
    if name in ("warp_stack_fwd", "warp_stack_bwd"):
        B, C, H, W, dt, lay = m["B"], m["C"], m["H"], m["W"], _dt(m["dtype"]), m["layout"]
        s = 4 if dt == torch.float32 else 2
        x = _t((B, C, 2, H, W), dev, dt, lay, gen)
        flow = torch.randn(B, 2, 2, H, W, device=dev, generator=gen) * sigma
        xs, ys = ops.base_grid(W, dev), ops.base_grid(H, dev)
        if name == "warp_stack_fwd":
            out = _t((B, C, 4, H, W), dev, dt, lay, gen)
            fn = lambda: _lib.check(lib.smow_warp_stack_fwd(x.data_ptr(), flow.data_ptr(), xs.data_ptr(), ys.data_ptr(),  # noqa: E731
                                                            out.data_ptr(), B, C, H, W, m["dtype"], lay, st()), name)
            return fn, ops.warp_fwd_bytes(B, C, H, W, s), ops.warp_fwd_bytes(B, C, H, W, s), (x, flow, out)
        gout = _t((B, C, 4, H, W), dev, dt, lay, gen)
        gx, gflow = torch.empty_like(x), torch.empty_like(flow)
        ws = torch.zeros(64, dtype=torch.uint8, device=dev)
        wsn = 64
        if lay == _lib.NDHWC and dt == torch.float32 and _lib.get_option("warp_bwd_variant") == 3:
            wsn = int(lib.smow_warp_bwd_workspace_bytes(B, H, W))
            ws = torch.empty(wsn, dtype=torch.uint8, device=dev)
        fn = lambda: _lib.check(lib.smow_warp_stack_bwd(gout.data_ptr(), x.data_ptr(), flow.data_ptr(), xs.data_ptr(),  # noqa: E731
                                                        ys.data_ptr(), gx.data_ptr(), gflow.data_ptr(), B, C, H, W, m["dtype"],
                                                        lay, ws.data_ptr(), wsn, st()), name)
        return fn, ops.warp_bwd_bytes(B, C, H, W, s), ops.warp_bwd_bytes(B, C, H, W, s), (x, flow, gout, gx, gflow, ws)
    if name in ("tlerp_cat_fwd", "tlerp_cat_bwd"):
        B, Cd, Cs, hw, dt, lay = m["B"], m["Cd"], m["Cs"], m["hw"], _dt(m["dtype"]), m["layout"]
        s = 4 if dt == torch.float32 else 2
        skip = _t((B, max(Cs, 4), 2, hw, 1), dev, dt, lay, gen) if Cs else None
        if Cs == 0 and m.get("act") != 2:
            raise KeyError("tlerp without a skip")
        skip = skip if Cs else torch.zeros(1, device=dev)
        cat = _t((B, Cd + Cs, 4, hw, 1), dev, dt, lay, gen)
        if m.get("act") == 2:  # BatchNorm-apply + LeakyReLU folded in (ops.bn_act_tlerp_cat): y -> cat[:, :Cd]
            z = _t((B, Cd, 4, hw, 1), dev, torch.float32, lay, gen)
            bn = torch.rand(6, Cd, device=dev, generator=gen) + 0.5
            s1 = skip if Cs else None
            s2 = skip[:, :, 1] if Cs else None
            if name == "tlerp_cat_fwd":
                fn = lambda: _lib.check(lib.smow_bn_act_tlerp_cat_fwd(z.data_ptr(), bn.data_ptr(), s1.data_ptr() if Cs else None,  # noqa: E731
                                                                      s2.data_ptr() if Cs else None, cat.data_ptr(), B, Cd, Cs, hw,
                                                                      2 * Cs * hw, 0.2, st()), name)
                alg = ops.tlerp_fwd_bytes(B, Cd, Cs, hw, 4) + ops.bn_act_fwd_bytes(B, Cd, hw)
                return fn, alg, alg, (z, skip, cat, bn)
            gy = torch.empty_like(z)
            gskip = torch.empty_like(skip) if Cs else None
            g2 = gskip[:, :, 1] if Cs else None
            dg, db = torch.empty(Cd, device=dev), torch.empty(Cd, device=dev)
            n = int(lib.smow_bn_act_bwd_workspace_bytes(B, Cd, hw))
            ws = torch.empty(max(n, 16), dtype=torch.uint8, device=dev)

            def fn():
                _lib.check(lib.smow_bn_act_bwd_reduce(cat.data_ptr(), z.data_ptr(), bn.data_ptr(), dg.data_ptr(), db.data_ptr(), B, Cd, Cs,
                                                      hw, 0.2, ws.data_ptr(), n, st()), name)
                _lib.check(lib.smow_bn_act_tlerp_cat_bwd(cat.data_ptr(), z.data_ptr(), bn.data_ptr(), gy.data_ptr(),
                                                         gskip.data_ptr() if Cs else None, g2.data_ptr() if Cs else None, B, Cd, Cs, hw,
                                                         2 * Cs * hw, 0.2, st()), name)
            alg = ops.tlerp_bwd_bytes(B, Cs, hw, 4) + ops.bn_act_bwd_bytes(B, Cd, hw)
            return fn, alg, alg, (z, cat, gy, gskip, bn, ws)
        if m.get("act"):       # LeakyReLU of the decoder block folded in: z -> cat[:, :Cd], no copy (ops.act_tlerp_*)
            z = _t((B, Cd, 4, hw, 1), dev, dt, lay, gen)
            s2 = skip[:, :, 1]
            if name == "tlerp_cat_fwd":
                fn = lambda: _lib.check(lib.smow_act_tlerp_cat_fwd(z.data_ptr(), skip.data_ptr(), s2.data_ptr(), cat.data_ptr(), B, Cd,  # noqa: E731
                                                                   Cs, hw, 2 * Cs * hw, 0.2, m["dtype"], st()), name)
                alg = ops.tlerp_fwd_bytes(B, Cd, Cs, hw, s) + ops.act_cat_bytes(B, Cd, hw, s)
                return fn, alg, alg, (z, skip, cat)
            gz, gskip = torch.empty_like(z), torch.empty_like(skip)
            g2 = gskip[:, :, 1]
            fn = lambda: _lib.check(lib.smow_act_tlerp_cat_bwd(cat.data_ptr(), z.data_ptr(), gz.data_ptr(), gskip.data_ptr(), g2.data_ptr(),  # noqa: E731
                                                               B, Cd, Cs, hw, 2 * Cs * hw, 0.2, m["dtype"], st()), name)
            alg = ops.tlerp_bwd_bytes(B, Cs, hw, s) + ops.act_cat_bwd_bytes(B, Cd, hw, s)
            return fn, alg, alg, (z, cat, gz, gskip)
        if name == "tlerp_cat_fwd":
            dec = _t((B, Cd, 4, hw, 1), dev, dt, lay, gen) if Cd and m.get("copy_dec", True) else None
            fn = lambda: _lib.check(lib.smow_tlerp_cat_fwd(None if dec is None else dec.data_ptr(), skip.data_ptr(),  # noqa: E731
                                                           cat.data_ptr(), B, Cd, Cs, hw, m["dtype"], lay, st()), name)
            alg = ops.tlerp_fwd_bytes(B, Cd, Cs, hw, s)
            return fn, alg, alg + (ops.tlerp_fwd_overhead_bytes(B, Cd, hw, s) if dec is not None else 0), (dec, skip, cat)
        gskip = torch.empty_like(skip)
        fn = lambda: _lib.check(lib.smow_tlerp_cat_bwd(cat.data_ptr(), gskip.data_ptr(), B, Cd, Cs, hw, m["dtype"], lay, st()), name)  # noqa: E731
        return fn, ops.tlerp_bwd_bytes(B, Cs, hw, s), ops.tlerp_bwd_bytes(B, Cs, hw, s), (cat, gskip)
    if name in ("tokenizer_fwd", "tokenizer_bwd"):
        B, C, hw = m["B"], m["C"], m["hw"]
        x = _t((B, C, 4, hw, 1), dev, torch.float32, _lib.NDHWC, gen)
        wa, ba = torch.randn(8, C, device=dev, generator=gen) / 4, torch.randn(8, device=dev, generator=gen)
        tokens, stats = torch.empty(B, 4, 8, C, device=dev), torch.empty(B, 4, 16, device=dev)
        n = int(lib.smow_tokenizer_workspace_bytes(B, C, hw))
        ws = torch.empty(max(n, 16), dtype=torch.uint8, device=dev)
        fwd = lambda: _lib.check(lib.smow_tokenizer_fwd(x.data_ptr(), wa.data_ptr(), ba.data_ptr(), tokens.data_ptr(),  # noqa: E731
                                                        stats.data_ptr(), B, C, hw, _lib.F32, _lib.NDHWC, ws.data_ptr(), n, st()), name)
        if name == "tokenizer_fwd":
            return fwd, ops.tokenizer_fwd_bytes(B, C, hw), ops.tokenizer_fwd_bytes(B, C, hw), (x, ws)
        fwd()                                                              # tokens / stats must be the forward's outputs
        gt = torch.randn(B, 4, 8, C, device=dev, generator=gen)
        gx, gwa, gba = torch.empty_like(x), torch.empty_like(wa), torch.empty_like(ba)
        fn = lambda: _lib.check(lib.smow_tokenizer_bwd(gt.data_ptr(), x.data_ptr(), wa.data_ptr(), ba.data_ptr(),  # noqa: E731
                                                       tokens.data_ptr(), stats.data_ptr(), gx.data_ptr(), gwa.data_ptr(),
                                                       gba.data_ptr(), B, C, hw, _lib.F32, _lib.NDHWC, ws.data_ptr(), n, st()), name)
        return fn, ops.tokenizer_bwd_bytes(B, C, hw), ops.tokenizer_bwd_bytes(B, C, hw), (x, gx, ws, gt)
    if name in ("warp_tokens_fwd", "warp_tokens_bwd"):
        B, C, H, W = m["B"], m["C"], m["H"], m["W"]
        x = _t((B, C, 2, H, W), dev, torch.float32, _lib.NDHWC, gen)
        flow = torch.randn(B, 2, 2, H, W, device=dev, generator=gen) * sigma
        xs, ys = ops.base_grid(W, dev), ops.base_grid(H, dev)
        wa, ba = torch.randn(8, C, device=dev, generator=gen) / 4, torch.randn(8, device=dev, generator=gen)
        tokens, stats = torch.empty(B, 4, 8, C, device=dev), torch.empty(B, 4, 16, device=dev)
        n = int(lib.smow_tokenizer_workspace_bytes(B, C, H * W))
        ws = torch.empty(max(n, 16), dtype=torch.uint8, device=dev)
        fwd = lambda: _lib.check(lib.smow_warp_tokenizer_fwd(x.data_ptr(), flow.data_ptr(), xs.data_ptr(), ys.data_ptr(),  # noqa: E731
                                                             wa.data_ptr(), ba.data_ptr(), tokens.data_ptr(), stats.data_ptr(),
                                                             B, C, H, W, _lib.F32, _lib.NDHWC, ws.data_ptr(), n, st()), name)
        if name == "warp_tokens_fwd":
            return fwd, ops.warp_tokens_fwd_bytes(B, C, H, W), ops.warp_tokens_fwd_bytes(B, C, H, W), (x, flow, ws)
        fwd()                                                              # tokens / stats must be the forward's outputs
        gt = torch.randn(B, 4, 8, C, device=dev, generator=gen)
        gstack = torch.empty((B, C, 4, H, W), device=dev).contiguous(memory_format=CL3)
        gwa, gba = torch.empty_like(wa), torch.empty_like(ba)
        fn = lambda: _lib.check(lib.smow_warp_tokenizer_bwd(gt.data_ptr(), x.data_ptr(), flow.data_ptr(), xs.data_ptr(),  # noqa: E731
                                                            ys.data_ptr(), wa.data_ptr(), ba.data_ptr(), tokens.data_ptr(),
                                                            stats.data_ptr(), gstack.data_ptr(), gwa.data_ptr(), gba.data_ptr(),
                                                            B, C, H, W, _lib.F32, _lib.NDHWC, ws.data_ptr(), n, st()), name)
        return fn, ops.warp_tokens_bwd_bytes(B, C, H, W), ops.warp_tokens_bwd_bytes(B, C, H, W), (x, flow, gstack, ws, gt)
    if name in ("frame_mix_fwd", "frame_mix_bwd", "frame_mix_wgrad"):
        B, C, T, hw, tc = m["B"], m["C"], m["T"], m["hw"], m.get("tc", 1)
        x = _t((B, C, T, hw, 1), dev, torch.float32, _lib.NDHWC, gen)
        y = _t((B, C, T, hw, 1), dev, torch.float32, _lib.NDHWC, gen)
        pack = torch.randn(1 + T, C, C, device=dev, generator=gen) / C ** 0.5
        if name != "frame_mix_wgrad":
            sh, off = (1, 1) if name == "frame_mix_fwd" else ((T - 1) % T, 0)
            if tc:
                fn = lambda: _lib.check(lib.smow_frame_mix_apply_tc(x.data_ptr(), pack.data_ptr(), None, y.data_ptr(), B, C, T,  # noqa: E731
                                                                    hw, C, sh, off, st()), name)
            else:
                fn = lambda: _lib.check(lib.smow_frame_mix_apply(x.data_ptr(), pack[0].data_ptr(), pack[1:].data_ptr(),  # noqa: E731
                                                                 y.data_ptr(), B, C, hw, sh, off, st()), name)
            nb = ops.frame_mix_apply_bytes(B, C, T, hw)
            return fn, nb, nb, (x, y, pack)
        gw = torch.empty(1 + T, C, C, device=dev)
        n = int(lib.smow_frame_mix_wgrad_tc_workspace_bytes(B, C, T, hw)) if tc else int(lib.smow_frame_mix_wgrad_workspace_bytes(B, C, hw))
        ws = torch.empty(max(n, 16), dtype=torch.uint8, device=dev)
        if tc:
            fn = lambda: _lib.check(lib.smow_frame_mix_wgrad_tc(x.data_ptr(), y.data_ptr(), gw.data_ptr(), B, C, T, hw, 1, 1,  # noqa: E731
                                                                ws.data_ptr(), n, st()), name)
        else:
            fn = lambda: _lib.check(lib.smow_frame_mix_wgrad(x.data_ptr(), y.data_ptr(), gw.data_ptr(), B, C, hw, ws.data_ptr(), n, st()), name)  # noqa: E731
        nb = ops.frame_mix_wgrad_bytes(B, C, T, hw)
        return fn, nb, nb, (x, y, gw, ws)
    if name in ("flow_head_fwd", "flow_head_bwd"):
        B, C, H, W, h, w = m["B"], m["C"], m["H"], m["W"], m["h"], m["w"]
        x = _t((B, C, 2, H, W), dev, torch.float32, _lib.NDHWC, gen)
        weight = torch.randn(2, 2 * C, 3, 3, 3, device=dev, generator=gen) / (36 * C) ** 0.5
        z = torch.randn(B, 9, h, w, 4, device=dev, generator=gen)
        flow = torch.randn(B, 2, 2, H, W, device=dev, generator=gen)
        n = int(lib.smow_flow_head_workspace_bytes(B, C, H, W))
        ws = torch.empty(n, dtype=torch.uint8, device=dev)
        if name == "flow_head_fwd":
            fn = lambda: _lib.check(lib.smow_flow_head_fwd(x.data_ptr(), weight.data_ptr(), z.data_ptr(), flow.data_ptr(), B, C, H, W,  # noqa: E731
                                                           h, w, ws.data_ptr(), n, st()), name)
            nb = ops.flow_head_fwd_bytes(B, C, H, W, h, w)
            return fn, nb, nb, (x, z, flow, ws)
        gx, gw, gz = torch.empty_like(x), torch.empty_like(weight), torch.empty_like(z)
        fn = lambda: _lib.check(lib.smow_flow_head_bwd(flow.data_ptr(), x.data_ptr(), weight.data_ptr(), gx.data_ptr(), gw.data_ptr(),  # noqa: E731
                                                       gz.data_ptr(), B, C, H, W, h, w, ws.data_ptr(), n, st()), name)
        nb = ops.flow_head_bwd_bytes(B, C, H, W, h, w)
        return fn, nb, nb, (x, z, flow, gx, gz, ws)
    raise KeyError(name)


def replay_ms(fns, reps=None, iters=5):
    """Median time per call of a CUDA graph that enqueues `fns` round-robin `reps` times."""
    reps = reps or max(len(fns), 8)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for f in fns:
            f()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fns[i % len(fns)]()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) / reps)
    ts.sort()
    del g
    return ts[len(ts) // 2]


def time_call(name, meta, dev, footprint=1 << 30, max_sets=64, sigma=0.3, seed=0):
    """-> dict(cold_ms, warm_ms, bytes, operand_bytes, sets).  `cold`: buffer sets rotate over >= footprint bytes."""
    gen = torch.Generator(device=dev).manual_seed(seed)
    first = build(name, meta, dev, gen, sigma)
    alg, opb = first[1], first[2]
    nsets = int(max(1, min(max_sets, -(-footprint // max(1, opb)))))
    sets = [first] + [build(name, meta, dev, gen, sigma) for _ in range(nsets - 1)]
    cold = replay_ms([s[0] for s in sets])
    warm = replay_ms([sets[0][0]], reps=8)
    del sets
    return {"cold_ms": cold, "warm_ms": warm, "bytes": alg, "operand_bytes": opb, "sets": nsets}
