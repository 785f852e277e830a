"""Semantic tokenizer, token transformers and the pixel-shuffle head (cuBLAS; not hot path).

Parameter names and arithmetic follow the reference (models/SMOW_Net.py:161-408; the LW file
repeats them at :180-427) so checkpoints load strictly; einops is replaced by plain
view/permute.  ``Transformer_Encoder`` is the sole consumer of the warped stack (row N2).
"""
import torch
import torch.nn as nn


def split_heads(t, heads):      # 'b n (h d) -> b h n d'
    b, n, hd = t.shape
    return t.view(b, n, heads, hd // heads).transpose(1, 2)


def merge_heads(t):             # 'b h n d -> b n (h d)'
    b, h, n, d = t.shape
    return t.transpose(1, 2).reshape(b, n, h * d)


class PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn

    def forward(self, x, **kw):
        return self.fn(self.norm(x), **kw)


class PreNorm2(nn.Module):
    """One LayerNorm applied to both the queries and the memory (models/SMOW_Net.py:326-334)."""

    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn

    def forward(self, x, m, **kw):
        return self.fn(self.norm(x), self.norm(m), **kw)


class Residual(nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x, *rest, **kw):
        return self.fn(x, *rest, **kw) + x


Residual2 = Residual  # the reference's two-input flavour has the same single attribute ``fn``


class FeedForward(nn.Module):
    def __init__(self, dim, hidden_dim, dropout=0.0):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden_dim, dim), nn.Dropout(dropout))

    def forward(self, x):
        return self.net(x)


class Attention(nn.Module):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        inner = dim_head * heads
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.attend = nn.Softmax(dim=-1)
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        if heads == 1 and dim_head == dim:
            self.to_out = nn.Identity()
        else:
            self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout))

    def forward(self, x):
        q, k, v = (split_heads(t, self.heads) for t in self.to_qkv(x).chunk(3, dim=-1))
        attn = self.attend(torch.matmul(q, k.transpose(-1, -2)) * self.scale)
        return self.to_out(merge_heads(torch.matmul(attn, v)))


class Transformer(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0.0):
        super().__init__()
        self.layers = nn.ModuleList([
            nn.ModuleList([PreNorm(dim, Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout)),
                           PreNorm(dim, FeedForward(dim, mlp_dim, dropout=dropout))])
            for _ in range(depth)])

    def forward(self, x):
        for attn, ff in self.layers:
            x = attn(x) + x
            x = ff(x) + x
        return x


class Transformer_Encoder(nn.Module):
    """Per-frame spatial-attention pooling of (B,C,4,H,W) to 8 tokens of 4C, then one transformer
    layer (models/SMOW_Net.py:161-190)."""

    def __init__(self, in_chan=32, token_len=8, heads=8):
        super().__init__()
        self.token_len = token_len
        self.conv_a = nn.Conv2d(in_chan, token_len, kernel_size=1, padding=0)
        self.pos_embedding = nn.Parameter(torch.randn(4, token_len, in_chan))
        self.transformer = Transformer(dim=in_chan * 4, depth=1, heads=heads, dim_head=in_chan * 4,
                                       mlp_dim=in_chan * 4, dropout=0)

    def from_warp(self, ofw, x):
        """``self(ofw(x))`` — the reference's ``Transformer_Encoder(OFW(x0))`` (models/SMOW_Net.py:50-52) — without the
        warped stack in HBM where the fused kernels exist (rows A1 + N2: ``ops.warp_tokens``)."""
        from .. import ops
        if self.token_len == 8 and ops.warp_tokens_supported(x, self.conv_a.weight):
            b, c = x.shape[:2]
            tok = ops.warp_tokens(x, ofw.predict_flow(x), self.conv_a.weight, self.conv_a.bias) + self.pos_embedding.unsqueeze(0)
            return self.transformer(tok.permute(0, 2, 1, 3).reshape(b, self.token_len, 4 * c))
        return self(ofw(x))

    def forward(self, x):
        b, c, t, h, w = x.shape
        assert t == 4, "The time dimension (t) must be 4."
        if x.is_cuda and x.dtype == torch.float32 and self.token_len == 8 and c % 4 == 0 and (c // 4) & (c // 4 - 1) == 0 \
                and c <= 128:
            # row N2: all four frames pooled by one hand-written pass over the stack (ops.semantic_tokens)
            from .. import ops
            tok = ops.semantic_tokens(x, self.conv_a.weight, self.conv_a.bias) + self.pos_embedding.unsqueeze(0)
            return self.transformer(tok.permute(0, 2, 1, 3).reshape(b, self.token_len, t * c))
        per_frame = []
        for k in range(t):
            frame = x[:, :, k]
            attn = torch.softmax(self.conv_a(frame).reshape(b, self.token_len, -1), dim=-1)
            tokens = torch.einsum("bln,bcn->blc", attn, frame.reshape(b, c, -1))
            per_frame.append(tokens + self.pos_embedding[k])
        return self.transformer(torch.cat(per_frame, dim=2))


class Cross_Attention(nn.Module):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0, softmax=True):
        super().__init__()
        inner = dim_head * heads          # the reference passes dim_head=True -> inner = heads
        self.heads = heads
        self.scale = dim ** -0.5
        self.softmax = softmax
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(dim, inner, bias=False)
        self.to_v = nn.Linear(dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout))

    def forward(self, x, m, mask=None):
        if mask is not None:
            raise NotImplementedError("masked cross-attention is never used by SMOW-Net")
        q, k, v = split_heads(self.to_q(x), self.heads), split_heads(self.to_k(m), self.heads), \
            split_heads(self.to_v(m), self.heads)
        dots = torch.einsum("bhid,bhjd->bhij", q, k) * self.scale
        attn = dots.softmax(dim=-1) if self.softmax else dots
        return self.to_out(merge_heads(torch.einsum("bhij,bhjd->bhid", attn, v)))


class TransformerDecoder(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout, softmax=True):
        super().__init__()
        self.layers = nn.ModuleList([
            nn.ModuleList([Residual2(PreNorm2(dim, Cross_Attention(dim, heads=heads, dim_head=dim_head,
                                                                   dropout=dropout, softmax=softmax))),
                           Residual(PreNorm(dim, FeedForward(dim, mlp_dim, dropout=dropout)))])
            for _ in range(depth)])

    def forward(self, x, m, mask=None):
        for attn, ff in self.layers:
            x = ff(attn(x, m, mask=mask))
        return x


class Transformer_Decoder(nn.Module):
    """Pixels of the folded (B, C*4, H, W) decoder output attend to the 8 tokens
    (models/SMOW_Net.py:270-282)."""

    def __init__(self, in_chan=128, heads=8):
        super().__init__()
        self.transformer_decoder = TransformerDecoder(dim=in_chan, depth=1, heads=heads, dim_head=True,
                                                      mlp_dim=in_chan * 2, dropout=0, softmax=in_chan)

    def forward(self, x, m):
        b, c, t, h, w = x.shape
        seq = x.reshape(b, c * t, h * w).transpose(1, 2)            # 'b c h w -> b (h w) c'
        seq = self.transformer_decoder(seq, m)
        return seq.transpose(1, 2).reshape(b, c * t, h, w)


class Classifier(nn.Module):
    """1x1 conv to scale^2 maps + the reference's hand-rolled pixel shuffle (models/SMOW_Net.py:384-408),
    whose sub-pixel order is transposed w.r.t. F.pixel_shuffle: channel k -> (dy, dx) = (k % s, k // s)."""

    def __init__(self, in_chan, n_class, scale=2, pad=0):
        super().__init__()
        self.conv1 = nn.Conv2d(in_chan, n_class * scale * scale, kernel_size=1, padding=pad, bias=False)
        self.scale = scale
        self.n_class = n_class

    def forward(self, x):
        y = self.conv1(x)
        n, _, h, w = y.shape
        s = self.scale
        if self.n_class != 1:
            raise NotImplementedError("SMOW-Net uses n_class=1")
        y = y.view(n, s, s, h, w)                    # (n, dx, dy, h, w)
        return y.permute(0, 3, 2, 4, 1).reshape(n, 1, h * s, w * s)
