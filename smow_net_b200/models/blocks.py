"""Decoder / projection blocks shared by SMOW_Net and SMOW_Net_LW (out of the hot path, cuDNN).

Parameter names follow the reference so its checkpoints load strictly
(models/SMOW_Net.py:104-159,411-423 and models/SMOW_Net_LW.py:102-178); the code is organised
around one idea the reference spells out four times: a spatial (1,k,k) convolution whose four
output frames are then mixed *cyclically* by 1x1x1 convolutions,

    out[j] = shared(frame[j]) + own[j+1](frame[(j+1) % 4]),

``shared`` being initialised to the identity and every ``own`` to zero.
"""
import torch
import torch.nn as nn


def _cyclic_frame_mix_composed(frames5d, shared, own):
    """The reference's own composition (models/SMOW_Net.py:121-139): slice, ten 1x1x1 convs, adds, concat."""
    parts = [frames5d[:, :, k:k + 1] for k in range(4)]
    kept = [shared(p) for p in parts]
    lent = [m(p) for m, p in zip(own, parts)]
    return torch.cat([kept[j] + lent[(j + 1) % 4] for j in range(4)], dim=2)


_SIDE_STREAMS = {}


class _Fork:
    """Run independent pieces of one operator on side streams.  Only while a CUDA graph is being captured: the branches
    become parallel nodes of the graph, so the small decoder levels (a 16-CTA GEMM each) overlap instead of queueing;
    in eager mode the launch loop is the bottleneck and the context does nothing."""

    def __init__(self, device, n):
        self.active = torch.cuda.is_current_stream_capturing()
        if not self.active:
            return
        key = device.index
        pool = _SIDE_STREAMS.setdefault(key, [])
        while len(pool) < n:
            pool.append(torch.cuda.Stream(device=device))
        self.pool, self.main, self.done = pool, torch.cuda.current_stream(device), []
        self.start = self.main.record_event()

    def branch(self, i):
        if not self.active:
            import contextlib
            return contextlib.nullcontext()
        s = self.pool[i]
        s.wait_event(self.start)
        fork = self

        class _Ctx:
            def __enter__(self):
                self.cm = torch.cuda.stream(s)
                self.cm.__enter__()

            def __exit__(self, *exc):
                self.cm.__exit__(*exc)
                fork.done.append(s.record_event())
        return _Ctx()

    def join(self):
        if self.active:
            for e in self.done:
                self.main.wait_event(e)


def _mix_matrix(conv):
    """(Cin, Cout) matrix of a 1x1x1 Conv3d (weight is Cout x Cin) or ConvTranspose3d (weight is Cin x Cout)."""
    w = conv.weight[:, :, 0, 0, 0]
    return w if isinstance(conv, nn.ConvTranspose3d) else w.t()


class _CyclicMixGemm(torch.autograd.Function):
    """Row N4: the frame mix as plain GEMMs on the channels-last tensor, no slicing copies, no concat.

    In NDHWC memory a frame of one pair is a dense (HW, C) matrix, so with X = (B, 4, HW, Cin):
        Y           = X @ W_shared                      one GEMM over all four frames
        Y[:, j]    += X[:, (j+1) % 4] @ W_own[j+1]      four strided-batched GEMMs accumulating in place
    and the backward is the transposed set.  The GEMMs follow torch.backends.cuda.matmul.allow_tf32 as the process set it
    (fp32 by default); the global flag is never toggled per call — autograd worker threads share it."""

    @staticmethod
    def forward(ctx, x5, w_shared, w0, w1, w2, w3, bias):
        B, Cin, T, H, W = x5.shape
        X = x5.permute(0, 2, 3, 4, 1).reshape(B, T, H * W, Cin)          # a view: x5 is channels_last_3d
        ws = (w0, w1, w2, w3)
        Cout = w_shared.shape[1]
        Y = torch.mm(X.view(-1, Cin), w_shared).view(B, T, H * W, Cout)
        fork = _Fork(x5.device, 4)                     # the four frames are independent
        for j in range(4):
            k = (j + 1) % 4
            with fork.branch(j):
                Y[:, j].baddbmm_(X[:, k], ws[k].unsqueeze(0).expand(B, Cin, Cout))
        fork.join()
        if bias is not None:                                                 # (4, Cout): shared bias + own[j+1] bias
            Y.add_(bias.view(1, T, 1, Cout))
        ctx.save_for_backward(X, w_shared, w0, w1, w2, w3)
        ctx.has_bias = bias is not None
        ctx.dims = (B, Cin, Cout, T, H, W)
        return Y.view(B, T, H, W, Cout).permute(0, 4, 1, 2, 3)

    @staticmethod
    def backward(ctx, gy5):
        X, w_shared, w0, w1, w2, w3 = ctx.saved_tensors
        B, Cin, Cout, T, H, W = ctx.dims
        ws = (w0, w1, w2, w3)
        G = gy5.contiguous(memory_format=torch.channels_last_3d).permute(0, 2, 3, 4, 1).reshape(B, T, H * W, Cout)
        G2 = G.view(-1, Cout)
        gX = torch.mm(G2, w_shared.t()).view(B, T, H * W, Cin)
        gws = [None] * 4
        fork = _Fork(gy5.device, 4)
        for j in range(4):
            k = (j + 1) % 4
            with fork.branch(j):
                gX[:, k].baddbmm_(G[:, j], ws[k].t().unsqueeze(0).expand(B, Cout, Cin))
                gws[k] = torch.bmm(X[:, k].transpose(1, 2), G[:, j]).sum(0)
        fork.join()
        # K = B*4*HW is huge and M = N = C tiny: one GEMM per (pair, frame) + a sum parallelises where cuBLAS's
        # single un-split GEMM does not (305 us -> ~40 us at the 128 x 128 level)
        g_shared = torch.bmm(X.view(B * T, H * W, Cin).transpose(1, 2), G.view(B * T, H * W, Cout)).sum(0)
        gbias = G.sum(dim=(0, 2)) if ctx.has_bias else None               # (4, Cout)
        return gX.view(B, T, H, W, Cin).permute(0, 4, 1, 2, 3), g_shared, gws[0], gws[1], gws[2], gws[3], gbias


def tensor_core_mix_enabled():
    """The tcgen05 frame-mix kernels compute TF32 products (fp32 accumulation): exactly what cuDNN does for the 1x1x1
    convolutions they replace when torch.backends.cudnn.allow_tf32 is True (PyTorch's default) — so that flag decides.
    With it off (strict fp32) the exact SIMT kernels / GEMMs run instead."""
    return bool(torch.backends.cudnn.allow_tf32)


def fused_bn_enabled(frames5d, shared, own, bn, cs=0):
    """The fully fused decoder block tail — frame mix with BatchNorm statistics from its epilogue, then BatchNorm-apply +
    LeakyReLU (+ temporal lerp + concat) in one pass — needs: training-mode batch statistics, the tensor-core mix
    (cudnn.allow_tf32), C <= 64 (the persistent kernel) and 16-byte channel vectors.  Tests switch it off to route the block
    through the oracle seams."""
    if not (bn.training and bn.track_running_stats and bn.affine and bn.momentum is not None and len(own) == 4):
        return False
    if not (frames5d.is_cuda and frames5d.dtype == torch.float32 and frames5d.dim() == 5 and frames5d.shape[2] == 4):
        return False
    from .. import ops
    c_out = shared.weight.shape[1] if isinstance(shared, nn.ConvTranspose3d) else shared.weight.shape[0]
    return tensor_core_mix_enabled() and cs % 4 == 0 and ops.frame_mix_tc_stats_supported(frames5d, c_out, 4)


def mix_bn_act(frames5d, shared, own, bn, slope, skip=None, skip_pair=None):
    """leaky_relu(bn(cyclic_frame_mix(frames))) [concatenated with the temporally up-sampled skip] on the fused kernels."""
    from .. import ops
    bias = None
    if shared.bias is not None:
        bias = torch.stack([shared.bias + own[(f + 1) % 4].bias for f in range(4)])
    pack = torch.stack([m.weight[:, :, 0, 0, 0] for m in [shared] + list(own)])
    y, parts = ops.frame_mix_tc_stats(frames5d, pack, bias, 4, 1, 1, nk=not isinstance(shared, nn.ConvTranspose3d))
    return ops.bn_act_tlerp_cat(y, parts, bn, slope, skip=skip, skip_pair=skip_pair)


def cyclic_frame_mix(frames5d, shared, own, shift=1, own_off=1):
    """out[:, :, f] = shared(frames[f]) + own[(f + own_off) % T](frames[(f + shift) % T]),  T = len(own).

    Decoder blocks (reference models/SMOW_Net.py:121-139): T = 4, shift = own_off = 1, own[k] acts on frame k.
    Encoder temporal exchange (reference models/SMOW_Net.py:460-473): T = 2, shift = 1, own_off = 0,
    own = [conv3d_time_3, conv3d_time_1]."""
    T = len(own)
    on_gpu = frames5d.is_cuda and frames5d.dim() == 5 and frames5d.shape[2] == T and frames5d.dtype == torch.float32
    if not on_gpu:
        return _frame_mix_composed(frames5d, shared, own, shift, own_off)
    from .. import ops
    c_out = shared.weight.shape[1] if isinstance(shared, nn.ConvTranspose3d) else shared.weight.shape[0]
    bias = None
    if shared.bias is not None:
        # frame f carries the shared bias plus the bias of the own-conv that feeds it
        bias = torch.stack([shared.bias + own[(f + own_off) % T].bias for f in range(T)])
    if tensor_core_mix_enabled() and ops.frame_mix_tc_supported(frames5d, c_out, T):
        # 1 + T matrices in the orientation the parameters are stored in: Conv3d (out, in) / ConvTranspose3d (in, out)
        pack = torch.stack([m.weight[:, :, 0, 0, 0] for m in [shared] + list(own)])
        return ops.frame_mix_tc(frames5d, pack, bias, T, shift, own_off, nk=not isinstance(shared, nn.ConvTranspose3d))
    if T != 4 or shift != 1 or own_off != 1 or not frames5d.is_contiguous(memory_format=torch.channels_last_3d):
        return _frame_mix_composed(frames5d, shared, own, shift, own_off)
    w_shared = _mix_matrix(shared)
    if ops.frame_mix_supported(frames5d, w_shared.shape[1]):
        # strict fp32, the large decoder levels (C = 16 / 28 / 32 / 64): exact SIMT kernels
        y = ops.frame_mix(frames5d, w_shared, torch.stack([_mix_matrix(m) for m in own]))
        return y if bias is None else y + bias.t().reshape(1, -1, 4, 1, 1)
    return _CyclicMixGemm.apply(frames5d, w_shared, *[_mix_matrix(m) for m in own], bias)


def _frame_mix_composed(frames5d, shared, own, shift, own_off):
    """The reference's composition, any T: slices, 1x1x1 convolutions, adds, concat."""
    T = len(own)
    parts = [frames5d[:, :, k:k + 1] for k in range(T)]
    return torch.cat([shared(parts[f]) + own[(f + own_off) % T](parts[(f + shift) % T]) for f in range(T)], dim=2)


def _identity_1x1(conv):
    with torch.no_grad():
        torch.nn.init.eye_(conv.weight[:, :, 0, 0, 0])


def _zero(conv):
    torch.nn.init.constant_(conv.weight, 0.0)


class TemporalDeconvMix(nn.Module):
    """x2 spatial transposed conv + cyclic frame mixing + BN + LeakyReLU(0.2).

    ``wide=True`` is SMOW_Net's block (5x5 kernel, biased time convs, models/SMOW_Net.py:104-139);
    ``wide=False`` is SMOW_Net_LW's (3x3 kernel, bias-free, models/SMOW_Net_LW.py:102-137)."""

    def __init__(self, in_dim, out_dim, wide=True):
        super().__init__()
        k = 5 if wide else 3
        self.conv3d_spatial = nn.ConvTranspose3d(in_dim, out_dim, kernel_size=(1, k, k), stride=(1, 2, 2),
                                                 padding=(0, k // 2, k // 2), output_padding=(0, 1, 1))
        t_in = in_dim if wide else out_dim
        for i in range(1, 6):
            setattr(self, "conv3d_time_%d" % i, nn.ConvTranspose3d(t_in, out_dim, kernel_size=1, bias=wide))
        for i in range(1, 5):
            _zero(getattr(self, "conv3d_time_%d" % i))
        _identity_1x1(self.conv3d_time_5)
        self.batch = nn.BatchNorm3d(out_dim)
        self.leaky = nn.LeakyReLU(0.2, inplace=True)

    def pre_activation(self, x):
        """Everything but the final LeakyReLU: transposed conv, frame mix, BatchNorm."""
        own = [getattr(self, "conv3d_time_%d" % i) for i in range(1, 5)]
        return self.batch(cyclic_frame_mix(self.conv3d_spatial(x), self.conv3d_time_5, own))

    def forward(self, x):
        return self.leaky(self.pre_activation(x))

    def forward_into_concat(self, x, skip=None, skip_pair=None):
        """forward(x) concatenated with the temporally up-sampled skip (rows A3 + A4).  On the GPU the block's LeakyReLU is
        folded into the launch that writes the concat buffer (ops.act_tlerp_*): no stand-alone activation pass, no copy of
        the decoder half.  `skip` (B,Cs,2,h,w) or `skip_pair` = two (B,Cs,h,w) frames."""
        from .. import ops
        cs = skip.shape[1] if skip is not None else skip_pair[0].shape[1]
        own = [getattr(self, "conv3d_time_%d" % i) for i in range(1, 5)]
        s = self.conv3d_spatial(x)
        if fused_bn_enabled(s, self.conv3d_time_5, own, self.batch, cs):
            # frame mix (+ BatchNorm statistics) -> BatchNorm-apply + LeakyReLU + lerp + concat: two passes over the block output
            return mix_bn_act(s, self.conv3d_time_5, own, self.batch, self.leaky.negative_slope, skip=skip, skip_pair=skip_pair)
        z = self.batch(cyclic_frame_mix(s, self.conv3d_time_5, own))
        if ops.act_cat_supported(z, cs):
            slope = self.leaky.negative_slope
            return ops.act_tlerp_cat(z, skip, slope) if skip is not None else ops.act_tlerp_pair_cat(z, *skip_pair, slope)
        y = self.leaky(z)
        return ops.tlerp_cat(y, skip) if skip is not None else ops.tlerp_pair_cat(y, *skip_pair)


class DoubleConv3d(nn.Module):
    """SMOW_Net's conv_block_2_3d: two 3x3x3 convs, BN after each, LeakyReLU between
    (models/SMOW_Net.py:142-156).  The Sequential keeps the reference's attribute name."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.conv_block_2_3d = nn.Sequential(
            nn.Conv3d(in_dim, out_dim, kernel_size=3, padding=1), nn.BatchNorm3d(out_dim),
            nn.LeakyReLU(0.2, inplace=True),
            nn.Conv3d(out_dim, out_dim, kernel_size=3, padding=1), nn.BatchNorm3d(out_dim))

    def forward(self, x):
        return self.conv_block_2_3d(x)


class SpatialConvMix(nn.Module):
    """SMOW_Net_LW's conv_block_2_3d: (1,3,3) conv + cyclic frame mixing + BN + LeakyReLU
    (models/SMOW_Net_LW.py:140-175)."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.conv3d_s = nn.Conv3d(in_dim, out_dim, kernel_size=(1, 3, 3), padding=(0, 1, 1))
        for i in range(1, 6):
            setattr(self, "conv3d_t%d" % i, nn.Conv3d(out_dim, out_dim, kernel_size=1, bias=False))
        for i in range(1, 5):
            _zero(getattr(self, "conv3d_t%d" % i))
        _identity_1x1(self.conv3d_t5)
        self.b = nn.BatchNorm3d(out_dim)
        self.l = nn.LeakyReLU(0.2, inplace=True)

    def forward(self, x):
        own = [getattr(self, "conv3d_t%d" % i) for i in range(1, 5)]
        s = self.conv3d_s(x)
        if fused_bn_enabled(s, self.conv3d_t5, own, self.b):
            return mix_bn_act(s, self.conv3d_t5, own, self.b, self.l.negative_slope)
        return self.l(self.b(cyclic_frame_mix(s, self.conv3d_t5, own)))


class PointwiseConvBN(nn.Module):
    """BasicConv3d: 1x1x1 conv + BN + ReLU under ``conv_bn`` (models/SMOW_Net.py:411-423)."""

    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv_bn = nn.Sequential(nn.Conv3d(in_ch, out_ch, kernel_size=1), nn.BatchNorm3d(out_ch), nn.ReLU())

    def forward(self, x):
        return self.conv_bn(x)


def convs_channels_last_3d(module):
    """Store every 5-D (3-D convolution) weight of `module` in channels_last_3d order, in place.  cuDNN then runs
    the decoder / encoder convolutions and BatchNorms as NDHWC kernels: 52.7 -> 42.7 ms (SMOW_Net) and
    38.1 -> 33.6 ms (SMOW_Net_LW) per batch-16 step on B200 (benchmarks/e2e_probe.py).  Shapes, names and values
    of the parameters are untouched, so state_dicts stay interchangeable with the reference."""
    for p in module.parameters():
        # 1x1x1 kernels are left alone: for them both memory formats describe the same bytes, and re-striding the
        # parameter only makes autograd's (contiguous-strided) gradient disagree with DDP's bucket view ("Grad strides
        # do not match bucket view strides": one extra copy per such gradient inside the reducer)
        if p.dim() == 5 and tuple(p.shape[2:]) != (1, 1, 1):
            p.data = p.data.contiguous(memory_format=torch.channels_last_3d)
    return module


def spatial_max_pool():
    """max_pooling_3d(): halves H and W, keeps the 4 frames (models/SMOW_Net.py:158-159)."""
    return nn.MaxPool3d(kernel_size=(1, 2, 2), stride=(1, 2, 2), padding=0)
