"""PyTorch restatement of the reference's hot path — TEST INFRASTRUCTURE ONLY.

These functions issue the same ATen calls, in the same order, as the reference code they
cite, so on any device they reproduce "the reference's own PyTorch path" (north_star) even
where /root/reference itself is not present (the GPU box).  oracle/make_golden.py asserts
bit-equality with the real reference in the build container.

The product package never imports this file.
"""
import torch
import torch.nn.functional as F


def ref_flow_warp(input, flow):
    """OFW.flow_warp — reference models/SMOW_Net.py:612-638 (= models/SMOW_Net_LW.py:454-480)."""
    B, C, T, H, W = input.shape
    # :617-621  CPU fp32 linspace base grid, shipped to the device each call
    ys = torch.linspace(-1.0, 1.0, H).view(-1, 1).expand(H, W)
    xs = torch.linspace(-1.0, 1.0, W).view(1, -1).expand(H, W)
    grid = torch.stack((xs, ys), 2).unsqueeze(0).expand(B, H, W, 2).type_as(input).to(input.device)
    norm = torch.tensor([[[[W, H]]]]).type_as(input).to(input.device)       # :627
    frames = []
    for t in range(T):                                                       # :623-632
        field = flow[:, :, t].permute(0, 2, 3, 1) / norm
        warped = F.grid_sample(input[:, :, t], (grid + field).clamp(-1, 1), mode="bilinear",
                               padding_mode="border", align_corners=True)
        frames.append(warped.unsqueeze(2))
    return torch.cat([input[:, :, 0:1]] + frames + [input[:, :, 1:2]], dim=2)   # :634-636


def ref_tlerp(skip):
    """Temporal 2 -> 4 upsample — reference models/SMOW_Net.py:64-73."""
    b, c, t, h, w = skip.shape
    return F.interpolate(skip, size=(4, h, w), mode="trilinear", align_corners=True)


def ref_tlerp_cat(dec, skip):
    """Upsample + decoder skip concat — reference models/SMOW_Net.py:64-73 and :78,82,86,90,94."""
    up = ref_tlerp(skip)
    return up if dec is None else torch.cat([dec, up], dim=1)


def ref_pair_stack(x_t1, x_t2):
    """Bi-temporal stacking — reference models/SMOW_Net_LW.py:38-40."""
    return torch.cat([x_t1.unsqueeze(2), x_t2.unsqueeze(2)], 2)


def warp_with_grads(x, flow, gout):
    """(out, grad_x, grad_flow) of ref_flow_warp by autograd, detached."""
    x = x.detach().clone().requires_grad_(True)
    flow = flow.detach().clone().requires_grad_(True)
    out = ref_flow_warp(x, flow)
    out.backward(gout)
    return out.detach(), x.grad.detach(), flow.grad.detach()


def tlerp_cat_with_grads(dec, skip, gcat):
    skip = skip.detach().clone().requires_grad_(True)
    d = None if dec is None else dec.detach().clone().requires_grad_(True)
    cat = ref_tlerp_cat(d, skip)
    cat.backward(gcat)
    return cat.detach(), (None if d is None else d.grad.detach()), skip.grad.detach()


def ref_semantic_tokens(x, weight, bias):
    """The reference's tokenizer loop (models/SMOW_Net.py:176-187 = models/SMOW_Net_LW.py:195-206), per frame:
    conv_a (1x1, C -> 8) -> view(b, 8, H*W) -> softmax(dim=-1) -> einsum('bln,bcn->blc').  Returns (B, 4, 8, C);
    the positional embedding and the concat of the four token sets are left to the caller."""
    b, c, t, h, w = x.shape
    out = []
    for k in range(t):
        frame = x[:, :, k]
        attn = torch.softmax(F.conv2d(frame, weight, bias).reshape(b, weight.shape[0], -1), dim=-1)
        out.append(torch.einsum("bln,bcn->blc", attn, frame.reshape(b, c, -1)))
    return torch.stack(out, 1)


def ref_cyclic_frame_mix(frames, w_shared, w_own, bias=None, shift=1, own_off=1):
    """Matrix restatement of the temporal frame mix.  Decoder blocks (reference models/SMOW_Net.py:121-139,
    models/SMOW_Net_LW.py:119-137,160-175), T = 4, shift = own_off = 1:  out[:, :, j] = W5 . T_j + W_{j+1} . T_{j+1}
    (indices mod 4), i.e. `T1_F1 + T2_F2, T2_F1 + T3_F2, T3_F1 + T4_F2, T4_F1 + T1_F2` with *_F1 = conv3d_time_5 and
    Tk_F2 = conv3d_time_k.  Encoder Decompose_conv (reference models/SMOW_Net.py:460-473), T = 2, shift = 1, own_off = 0,
    w_own = [time_3, time_1]:  out[T1] = time_2(T1) + time_3(T2),  out[T2] = time_2(T2) + time_1(T1).
    frames (B,Cin,T,H,W); w_shared (Cin,Cout) and w_own (T,Cin,Cout) with rows = input channels (a Conv3d weight
    transposed, a ConvTranspose3d weight as stored); bias (T,Cout) = shared bias + bias of the own conv feeding frame f."""
    T = w_own.shape[0]
    out = []
    for f in range(T):
        g, k = (f + own_off) % T, (f + shift) % T
        y = torch.einsum("bchw,cd->bdhw", frames[:, :, f], w_shared) + torch.einsum("bchw,cd->bdhw", frames[:, :, k], w_own[g])
        if bias is not None:
            y = y + bias[f].view(1, -1, 1, 1)
        out.append(y)
    return torch.stack(out, 2)


def ref_flow_head(x, coarse, weight):
    """The OFW flow head — reference models/SMOW_Net.py:606-608 (= models/SMOW_Net_LW.py:448-450), with the hard-coded
    (2,128,128) target spelled as the size of x (the same thing at 256x256 network inputs):
        seg_down = F.interpolate(seg_down, size=(2,H,W), mode='trilinear', align_corners=True)
        flow = self.flow_make(torch.cat([x, seg_down], dim=1))          # Conv3d(2C, 2, 3, padding=1, bias=False)
    x (B,C,2,H,W); coarse (B,C,2,h,w) = down(x); weight (2,2C,3,3,3) -> flow (B,2,2,H,W)."""
    up = F.interpolate(coarse, size=tuple(x.shape[2:]), mode="trilinear", align_corners=True)
    return F.conv3d(torch.cat([x, up], dim=1), weight, None, 1, 1)
