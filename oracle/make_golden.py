"""Generate tests/golden/*.npz by running the REAL reference (read-only tree at /root/reference).

TEST INFRASTRUCTURE ONLY.  Run in the build container (the GPU box has no /root/reference):

    python -m oracle.make_golden            # writes tests/golden/hotpath_ops.npz (+ model fixtures)

What it pins:
  * OFW.flow_warp forward + autograd backward      (reference models/SMOW_Net.py:612-638)
  * F.interpolate(...,(4,h,w),trilinear)+torch.cat  (reference models/SMOW_Net.py:64-73,78-94)
  * oracle.torch_ref restatements == reference, bit for bit, on the same inputs
  * oracle C restatement within 2e-6 of the reference
  * whole-model outputs of SMOW_Net / SMOW_Net_LW for seeded weights (see model_fixture)
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("SMOW_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# (name, B, C, H, W, flow sigma, flow offset) — small enough for pure CPU, shaped to hit every branch
WARP_CASES = [
    ("init_like", 2, 8, 16, 24, 0.3, 0.0),       # sub-pixel flows, like a freshly initialised OFW
    ("stress", 1, 4, 12, 12, 8.0, 0.0),          # +-4 px, clamp and border masks active
    ("zero_flow", 2, 4, 12, 20, 0.0, 0.0),       # the reference is NOT an identity here
    ("odd_shape", 1, 3, 5, 7, 2.0, 0.0),         # W, H not multiples of anything
    ("all_clamped", 1, 2, 8, 8, 0.5, 400.0),     # every coordinate clamped to +1: gradients gated to 0
    ("model_width", 1, 4, 24, 128, 0.3, 0.0),   # the models' row width (W=128), a 24-row strip
]
TLERP_CASES = [("with_dec", 2, 5, 6, 4, 4), ("no_dec", 1, 0, 3, 2, 8), ("x4_like", 1, 8, 16, 8, 8)]


def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit("reference tree %s not found: golden vectors can only be generated in the build container" % REF)
    sys.path.insert(0, REF)
    import models.SMOW_Net as ref_s  # noqa: E402
    return ref_s


def gen_ops(ref_s):
    from oracle import c_oracle, torch_ref
    out = {}
    ofw = ref_s.OFW(4)  # flow_warp uses no parameters; any inplane works
    g = torch.Generator().manual_seed(2022)
    for name, B, C, H, W, sigma, off in WARP_CASES:
        x = torch.randn(B, C, 2, H, W, generator=g)
        flow = torch.randn(B, 2, 2, H, W, generator=g) * sigma + off
        gout = torch.randn(B, C, 4, H, W, generator=g)
        xr = x.clone().requires_grad_(True)
        fr = flow.clone().requires_grad_(True)
        y = ofw.flow_warp(xr, fr, (H, W))
        y.backward(gout)
        # restatement must be the same computation, bit for bit
        y2, gx2, gf2 = torch_ref.warp_with_grads(x, flow, gout)
        assert torch.equal(y, y2) and torch.equal(xr.grad, gx2) and torch.equal(fr.grad, gf2), name
        # C oracle within fp32 rounding
        xs, ys = torch.linspace(-1, 1, W).numpy(), torch.linspace(-1, 1, H).numpy()
        assert np.array_equal(c_oracle.linspace_table(W), xs) and np.array_equal(c_oracle.linspace_table(H), ys), name
        yc = c_oracle.warp_stack_fwd(x.numpy(), flow.numpy(), xs, ys)
        gxc, gfc = c_oracle.warp_stack_bwd(gout.numpy(), x.numpy(), flow.numpy(), xs, ys)
        e = [np.abs(yc - y.detach().numpy()).max(), np.abs(gxc - xr.grad.numpy()).max(),
             np.abs(gfc - fr.grad.numpy()).max()]
        scale = max(1.0, float(np.abs(fr.grad.numpy()).max()))
        print("warp %-12s C-oracle vs reference: out %.2e  gx %.2e  gflow %.2e (|gflow|max %.1f)" % (name, *e, scale))
        assert e[0] < 2e-6 and e[1] < 4e-6 and e[2] < 4e-6 * scale, name
        for k, v in (("x", x), ("flow", flow), ("gout", gout), ("out", y.detach()), ("gx", xr.grad), ("gflow", fr.grad)):
            out["warp/%s/%s" % (name, k)] = v.numpy().astype(np.float32)
    import torch.nn.functional as F
    for name, B, Cd, Cs, h, w in TLERP_CASES:
        skip = torch.randn(B, Cs, 2, h, w, generator=g)
        dec = torch.randn(B, Cd, 4, h, w, generator=g) if Cd else None
        gcat = torch.randn(B, Cd + Cs, 4, h, w, generator=g)
        sr = skip.clone().requires_grad_(True)
        dr = dec.clone().requires_grad_(True) if Cd else None
        up = F.interpolate(sr, size=(4, h, w), mode="trilinear", align_corners=True)   # reference :65
        cat = torch.cat([dr, up], dim=1) if Cd else up                                  # reference :94
        cat.backward(gcat)
        c2, gd2, gs2 = torch_ref.tlerp_cat_with_grads(dec, skip, gcat)
        assert torch.equal(cat, c2) and torch.equal(sr.grad, gs2), name
        cc = c_oracle.tlerp_cat_fwd(None if dec is None else dec.numpy(), skip.numpy())
        gsc = c_oracle.tlerp_cat_bwd(gcat.numpy(), Cd)
        e = [np.abs(cc - cat.detach().numpy()).max(), np.abs(gsc - sr.grad.numpy()).max()]
        print("tlerp %-10s C-oracle vs reference: cat %.2e  gskip %.2e" % (name, *e))
        assert e[0] < 1e-6 and e[1] < 2e-6, name
        for k, v in (("skip", skip), ("gcat", gcat), ("cat", cat.detach()), ("gskip", sr.grad)):
            out["tlerp/%s/%s" % (name, k)] = v.numpy().astype(np.float32)
        if Cd:
            out["tlerp/%s/dec" % name] = dec.numpy().astype(np.float32)
            out["tlerp/%s/gdec" % name] = dr.grad.numpy().astype(np.float32)
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, "hotpath_ops.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KiB" % (os.path.getsize(path) / 1024))


def gen_flow_head(ref_s):
    """Row N1: the flow head as the REFERENCE computes it — OFW.forward (models/SMOW_Net.py:604-610) run unmodified with
    its flow_warp replaced by a recorder, so the golden `flow` is the tensor the reference's own three lines
    (down, F.interpolate(..., (2,128,128)), flow_make(cat)) produced.  The oracle restatement must equal it bit for
    bit; its autograd gradients (x, coarse, weight as independent leaves) are stored beside it."""
    from oracle import torch_ref
    out = {}
    g = torch.Generator().manual_seed(606)
    for name, B, C in (("c2", 1, 2),):
        torch.manual_seed(608 + C)
        ofw = ref_s.OFW(C).train()
        with torch.no_grad():
            for p in ofw.parameters():
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
        x = torch.randn(B, C, 2, 128, 128, generator=g)         # the reference hard-codes the (2,128,128) target
        seen = {}
        ofw.flow_warp = lambda input, flow, size: seen.setdefault("flow", flow)      # recorder instead of the warp
        ofw(x)
        ref_flow = seen["flow"].detach()
        coarse = ofw.down(x).detach()                           # train-mode BN: same batch statistics as inside forward
        weight = ofw.flow_make.weight.detach()
        xr, cr, wr = (t.clone().requires_grad_(True) for t in (x, coarse, weight))
        mine = torch_ref.ref_flow_head(xr, cr, wr)
        assert torch.equal(mine.detach(), ref_flow), name       # restatement == reference, bit for bit
        gflow = torch.randn(mine.shape, generator=g)
        mine.backward(gflow)
        for k, v in (("x", x), ("coarse", coarse), ("weight", weight), ("gflow", gflow), ("flow", ref_flow),
                     ("gx", xr.grad), ("gcoarse", cr.grad), ("gweight", wr.grad)):
            out["flow_head/%s/%s" % (name, k)] = v.numpy().astype(np.float32)
        print("flow_head %-6s restatement == reference OFW.forward (|flow| max %.3f)" % (name, float(ref_flow.abs().max())))
    path = os.path.join(GOLDEN, "flow_head.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KiB" % (os.path.getsize(path) / 1024))


def main():
    torch.manual_seed(2022)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref_s = import_reference()
    if "--flow-head-only" in sys.argv:
        return gen_flow_head(ref_s)
    gen_ops(ref_s)
    gen_flow_head(ref_s)
    if "--ops-only" not in sys.argv:
        from oracle import model_fixture
        model_fixture.generate(REF, GOLDEN)


if __name__ == "__main__":
    main()
