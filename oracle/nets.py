"""Oracle networks: plain torch.nn fp32 NCDHW restatements of the two MONAI nets the
reference instantiates by name (3d_ldm/config/config_train_32g.json:8,41 via
3d_ldm/utils.py:243-246).  TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.

Deliberately un-optimised: F.conv3d / F.group_norm / F.silu / F.linear / softmax.
Module/attribute names are chosen so that state_dict() keys equal the MONAI (>=1.4)
key names listed in SURVEY.md Appendix E; a MONAI-trained checkpoint loaded by
3d_ldm/inference.py:72-77 must load here unchanged.
"""
from __future__ import annotations

import math
from typing import Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

# upper bound on one materialised block of attention scores (see SelfAttention.forward)
ATTN_SCORE_BYTES = 1 << 30

# --------------------------------------------------------------------------------------
# small building blocks
# --------------------------------------------------------------------------------------
class ConvOnly(nn.Module):
    """MONAI `Convolution(..., conv_only=True)`: a container whose single child is `.conv`
    (hence `*.conv.weight` keys)."""

    def __init__(self, cin: int, cout: int, kernel_size: int = 3, stride: int = 1, padding: int = 1):
        super().__init__()
        self.conv = nn.Conv3d(cin, cout, kernel_size, stride=stride, padding=padding)

    def forward(self, x):
        return self.conv(x)


def zero_module(m: nn.Module) -> nn.Module:
    for p in m.parameters():
        nn.init.zeros_(p)
    return m


class SelfAttention(nn.Module):
    """MONAI SABlock with separate q/k/v linears (use_combined_linear=False), qkv_bias=True,
    out_proj (include_fc=True), dropout 0.  Input [B, N, C]."""

    def __init__(self, hidden: int, heads: int):
        super().__init__()
        if hidden % heads:
            raise ValueError("hidden size should be divisible by num_heads.")
        self.heads = heads
        self.head_dim = hidden // heads
        self.scale = self.head_dim ** -0.5
        self.to_q = nn.Linear(hidden, hidden, bias=True)
        self.to_k = nn.Linear(hidden, hidden, bias=True)
        self.to_v = nn.Linear(hidden, hidden, bias=True)
        self.out_proj = nn.Linear(hidden, hidden)

    def forward(self, x):
        b, n, c = x.shape
        h, d = self.heads, self.head_dim

        def split(t):
            return t.view(b, n, h, d).permute(0, 2, 1, 3)  # [B,H,N,d]

        q, k, v = split(self.to_q(x)), split(self.to_k(x)), split(self.to_v(x))
        # the softmax is per query row, so evaluating the rows in blocks is the same arithmetic; it only bounds the
        # [B,H,rows,N] score matrix (64000 tokens x 64000 keys x 4 B = 16 GB in the 160^3 decoder) to ~1 GB
        rows = max(1, min(n, ATTN_SCORE_BYTES // max(1, 4 * b * h * n)))
        ys = []
        for r0 in range(0, n, rows):
            att = torch.einsum("bhxd,bhyd->bhxy", q[:, :, r0:r0 + rows], k) * self.scale
            att = att.softmax(dim=-1)
            ys.append(torch.einsum("bhxy,bhyd->bhxd", att, v))
        y = ys[0] if len(ys) == 1 else torch.cat(ys, dim=2)
        y = y.permute(0, 2, 1, 3).reshape(b, n, c)
        return self.out_proj(y)


class SpatialAttentionBlock(nn.Module):
    """GN -> flatten D*H*W tokens -> multi-head self-attention -> + residual (App. A)."""

    def __init__(self, channels: int, num_head_channels: int | None, groups: int, eps: float):
        super().__init__()
        self.norm = nn.GroupNorm(groups, channels, eps=eps, affine=True)
        heads = channels // num_head_channels if num_head_channels else 1
        self.attn = SelfAttention(channels, heads)

    def forward(self, x):
        res = x
        shp = x.shape
        t = self.norm(x).reshape(shp[0], shp[1], -1).transpose(1, 2)  # [B,N,C]
        t = self.attn(t)
        t = t.transpose(1, 2).reshape(shp)
        return t + res


class _Attention(nn.Module):
    """Shared core of MONAI SABlock / CrossAttentionBlock as DiffusionUNetTransformerBlock builds them
    (qkv_bias=False, include_fc=True, separate linears): q from x, k/v from `context` (x itself when None)."""

    def __init__(self, hidden: int, heads: int, head_dim: int, context_dim: int | None = None):
        super().__init__()
        inner = heads * head_dim
        self.heads, self.head_dim, self.scale = heads, head_dim, head_dim ** -0.5
        self.to_q = nn.Linear(hidden, inner, bias=False)
        self.to_k = nn.Linear(context_dim or hidden, inner, bias=False)
        self.to_v = nn.Linear(context_dim or hidden, inner, bias=False)
        self.out_proj = nn.Linear(inner, hidden)

    def forward(self, x, context=None):
        ctx = x if context is None else context
        b, n, _ = x.shape
        h, d = self.heads, self.head_dim
        q = self.to_q(x).view(b, n, h, d).permute(0, 2, 1, 3)
        k = self.to_k(ctx).view(b, ctx.shape[1], h, d).permute(0, 2, 1, 3)
        v = self.to_v(ctx).view(b, ctx.shape[1], h, d).permute(0, 2, 1, 3)
        att = (torch.einsum("bhxd,bhyd->bhxy", q, k) * self.scale).softmax(dim=-1)
        y = torch.einsum("bhxy,bhyd->bhxd", att, v).permute(0, 2, 1, 3).reshape(b, n, h * d)
        return self.out_proj(y)


class _GEGLUMLP(nn.Module):
    """MONAI MLPBlock(act="GEGLU"): linear1 -> (value, gate) halves -> value * gelu(gate) -> linear2 (dropout 0)."""

    def __init__(self, hidden: int, mlp_dim: int):
        super().__init__()
        self.linear1 = nn.Linear(hidden, mlp_dim * 2)
        self.linear2 = nn.Linear(mlp_dim, hidden)

    def forward(self, x):
        val, gate = self.linear1(x).chunk(2, dim=-1)
        return self.linear2(val * F.gelu(gate))


class TransformerBlock(nn.Module):
    """MONAI DiffusionUNetTransformerBlock: self attention, cross attention, GEGLU feed-forward, each pre-LayerNormed
    and residual.  [MONAI-recall >= 1.4; the reference only has the commented-out call, 3d_ldm/train_diffusion.py:314-321]"""

    def __init__(self, channels: int, heads: int, head_dim: int, cross_attention_dim: int | None):
        super().__init__()
        self.attn1 = _Attention(channels, heads, head_dim)
        self.ff = _GEGLUMLP(channels, channels * 4)
        self.attn2 = _Attention(channels, heads, head_dim, cross_attention_dim)
        self.norm1 = nn.LayerNorm(channels)
        self.norm2 = nn.LayerNorm(channels)
        self.norm3 = nn.LayerNorm(channels)

    def forward(self, x, context=None):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), context=context) + x
        return self.ff(self.norm3(x)) + x


class SpatialTransformer(nn.Module):
    """MONAI SpatialTransformer: GN -> 1x1x1 proj_in -> tokens -> transformer blocks -> 1x1x1 proj_out (zero-init)
    -> + residual.  Replaces SpatialAttentionBlock at attention levels when with_conditioning=True."""

    def __init__(self, channels: int, heads: int, head_dim: int, num_layers: int, groups: int, eps: float,
                 cross_attention_dim: int | None):
        super().__init__()
        inner = heads * head_dim
        self.norm = nn.GroupNorm(groups, channels, eps=eps, affine=True)
        self.proj_in = ConvOnly(channels, inner, 1, 1, 0)
        self.transformer_blocks = nn.ModuleList(
            [TransformerBlock(inner, heads, head_dim, cross_attention_dim) for _ in range(num_layers)])
        self.proj_out = zero_module(ConvOnly(inner, channels, 1, 1, 0))

    def forward(self, x, context=None):
        res = x
        b, _, d, h, w = x.shape
        t = self.proj_in(self.norm(x))
        inner = t.shape[1]
        t = t.permute(0, 2, 3, 4, 1).reshape(b, d * h * w, inner)
        for blk in self.transformer_blocks:
            t = blk(t, context=context)
        t = t.reshape(b, d, h, w, inner).permute(0, 4, 1, 2, 3).contiguous()
        return self.proj_out(t) + res


def _make_attn(cout, head_ch, groups, eps, xattn):
    """xattn = None: plain SpatialAttentionBlock; else (num_layers, cross_attention_dim): SpatialTransformer."""
    if xattn is None:
        return SpatialAttentionBlock(cout, head_ch, groups, eps)
    return SpatialTransformer(cout, cout // head_ch, head_ch, xattn[0], groups, eps, xattn[1])


class _NearestUp(nn.Module):
    """MONAI UpSample(mode="nontrainable", interp_mode="nearest", scale 2) + post conv;
    children named `upsample_non_trainable` (no params) and `postconv`."""

    def __init__(self, channels: int):
        super().__init__()
        self.upsample_non_trainable = nn.Upsample(scale_factor=2.0, mode="nearest")
        self.postconv = ConvOnly(channels, channels, 3, 1, 1)

    def forward(self, x):
        return self.postconv(self.upsample_non_trainable(x))


# --------------------------------------------------------------------------------------
# DiffusionModelUNet  
# --------------------------------------------------------------------------------------
def get_timestep_embedding(timesteps: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    if timesteps.ndim != 1:
        raise ValueError("Timesteps should be a 1d-array")
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half)
    args = timesteps[:, None].float() * freqs[None, :]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)  # cos first
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1))
    return emb


class UNetResBlock(nn.Module):
    def __init__(self, cin: int, cout: int, temb_ch: int, groups: int, eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps, affine=True)
        self.conv1 = ConvOnly(cin, cout, 3, 1, 1)
        self.time_emb_proj = nn.Linear(temb_ch, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps, affine=True)
        self.conv2 = zero_module(ConvOnly(cout, cout, 3, 1, 1))
        self.skip_connection = nn.Identity() if cin == cout else ConvOnly(cin, cout, 1, 1, 0)

    def forward(self, x, emb):
        h = self.conv1(F.silu(self.norm1(x)))
        t = self.time_emb_proj(F.silu(emb))[:, :, None, None, None]
        h = h + t
        h = self.conv2(F.silu(self.norm2(h)))
        return self.skip_connection(x) + h


class _DownOp(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.op = ConvOnly(ch, ch, 3, 2, 1)

    def forward(self, x):
        return self.op(x)


class _DownBlock(nn.Module):
    def __init__(self, cin, cout, temb, n_res, groups, eps, add_down, attn, head_ch, xattn=None):
        super().__init__()
        self.resnets = nn.ModuleList(
            [UNetResBlock(cin if j == 0 else cout, cout, temb, groups, eps) for j in range(n_res)]
        )
        if attn:
            self.attentions = nn.ModuleList([_make_attn(cout, head_ch, groups, eps, xattn) for _ in range(n_res)])
        self.has_attn, self.xattn = attn, xattn is not None
        self.downsampler = _DownOp(cout) if add_down else None

    def forward(self, h, emb, context=None):
        outs = []
        for j, r in enumerate(self.resnets):
            h = r(h, emb)
            if self.has_attn:
                h = (self.attentions[j](h, context) if self.xattn else self.attentions[j](h)).contiguous()
            outs.append(h)
        if self.downsampler is not None:
            h = self.downsampler(h)
            outs.append(h)
        return h, outs


class _MidBlock(nn.Module):
    def __init__(self, ch, temb, groups, eps, head_ch, xattn=None):
        super().__init__()
        self.resnet_1 = UNetResBlock(ch, ch, temb, groups, eps)
        self.attention = _make_attn(ch, head_ch, groups, eps, xattn)
        self.resnet_2 = UNetResBlock(ch, ch, temb, groups, eps)
        self.xattn = xattn is not None

    def forward(self, h, emb, context=None):
        h = self.resnet_1(h, emb)
        h = self.attention(h, context) if self.xattn else self.attention(h)
        return self.resnet_2(h, emb)


class _UpBlock(nn.Module):
    def __init__(self, cin, prev, cout, temb, n_res, groups, eps, add_up, attn, head_ch, xattn=None):
        super().__init__()
        blocks = []
        for j in range(n_res):
            skip_ch = cin if j == n_res - 1 else cout
            res_in = prev if j == 0 else cout
            blocks.append(UNetResBlock(res_in + skip_ch, cout, temb, groups, eps))
        self.resnets = nn.ModuleList(blocks)
        if attn:
            self.attentions = nn.ModuleList([_make_attn(cout, head_ch, groups, eps, xattn) for _ in range(n_res)])
        self.has_attn, self.xattn = attn, xattn is not None
        # MONAI >= 1.4: WrappedUpsample(UpSample) - an nn.Sequential with children `upsample_non_trainable` and `postconv` -
        # is assigned to `upsampler` directly, so the conv's key is `up_blocks.{i}.upsampler.postconv.conv.*`
        self.upsampler = _NearestUp(cout) if add_up else None

    def forward(self, h, skips, emb, context=None):
        for j, r in enumerate(self.resnets):
            h = torch.cat([h, skips.pop()], dim=1)
            h = r(h, emb)
            if self.has_attn:
                h = (self.attentions[j](h, context) if self.xattn else self.attentions[j](h)).contiguous()
        if self.upsampler is not None:
            h = self.upsampler(h)
        return h


def _per_level(v, n):
    return list(v) if isinstance(v, (list, tuple)) else [v] * n


def _alias_upsampler_keys(state_dict, prefix, *_):
    """load_state_dict pre-hook: the post-upsample conv of a UNet up block is `up_blocks.{i}.upsampler.postconv.conv.*`
    in MONAI >= 1.4 (WrappedUpsample subclasses UpSample and is assigned to `upsampler` itself); the nested spelling
    `upsampler.upsample.postconv.conv.*` of SURVEY.md App. E is accepted as an alias, so either checkpoint loads."""
    for k in [k for k in state_dict if k.startswith(prefix) and ".upsampler.upsample." in k]:
        state_dict[k.replace(".upsampler.upsample.", ".upsampler.")] = state_dict.pop(k)


class DiffusionModelUNet(nn.Module):
    def __init__(
        self,
        spatial_dims: int,
        in_channels: int,
        out_channels: int,
        num_res_blocks: Sequence[int] | int = (2, 2, 2, 2),
        channels: Sequence[int] = (32, 64, 64, 64),
        attention_levels: Sequence[bool] = (False, False, True, True),
        norm_num_groups: int = 32,
        norm_eps: float = 1e-6,
        resblock_updown: bool = False,
        num_head_channels: int | Sequence[int] = 8,
        with_conditioning: bool = False,
        transformer_num_layers: int = 1,
        cross_attention_dim: int | None = None,
        num_class_embeds: int | None = None,
        upcast_attention: bool = False,
        dropout_cattn: float = 0.0,
        include_fc: bool = True,
        use_combined_linear: bool = False,
        use_flash_attention: bool = False,
    ):
        super().__init__()
        if spatial_dims != 3:
            raise ValueError("oracle restates the 3-D path only")
        if resblock_updown:
            raise NotImplementedError("not used by any reference config (SURVEY.md 8f)")
        if with_conditioning and cross_attention_dim is None:
            raise ValueError("DiffusionModelUNet expects dimension of the cross-attention conditioning "
                             "(cross_attention_dim) when using with_conditioning.")
        if cross_attention_dim is not None and not with_conditioning:
            raise ValueError("DiffusionModelUNet expects with_conditioning=True when specifying the cross_attention_dim.")
        if any(c % norm_num_groups for c in channels):
            raise ValueError("DiffusionModelUNet expects all channels being multiple of norm_num_groups")
        if len(channels) != len(attention_levels):
            raise ValueError("DiffusionModelUNet expects channels being same size of attention_levels")
        L = len(channels)
        head = _per_level(num_head_channels, L)
        nres = _per_level(num_res_blocks, L)
        if len(head) != L or len(nres) != L:
            raise ValueError("per-level arguments must match len(channels)")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.block_out_channels = list(channels)
        self.with_conditioning = with_conditioning
        self._register_load_state_dict_pre_hook(_alias_upsampler_keys)
        temb = channels[0] * 4
        xattn = (transformer_num_layers, cross_attention_dim) if with_conditioning else None

        self.conv_in = ConvOnly(in_channels, channels[0], 3, 1, 1)
        self.time_embed = nn.Sequential(nn.Linear(channels[0], temb), nn.SiLU(), nn.Linear(temb, temb))
        self.num_class_embeds = num_class_embeds
        if num_class_embeds is not None:
            self.class_embedding = nn.Embedding(num_class_embeds, temb)

        self.down_blocks = nn.ModuleList()
        cout = channels[0]
        for i in range(L):
            cin, cout = cout, channels[i]
            self.down_blocks.append(
                _DownBlock(cin, cout, temb, nres[i], norm_num_groups, norm_eps, i != L - 1,
                           bool(attention_levels[i]), head[i], xattn)
            )
        self.middle_block = _MidBlock(channels[-1], temb, norm_num_groups, norm_eps, head[-1], xattn)

        self.up_blocks = nn.ModuleList()
        rc, rn = list(reversed(channels)), list(reversed(nres))
        ra, rh = list(reversed(attention_levels)), list(reversed(head))
        cout = rc[0]
        for i in range(L):
            prev, cout = cout, rc[i]
            cin = rc[min(i + 1, L - 1)]
            self.up_blocks.append(
                _UpBlock(cin, prev, cout, temb, rn[i] + 1, norm_num_groups, norm_eps, i != L - 1,
                         bool(ra[i]), rh[i], xattn)
            )
        self.out = nn.Sequential(
            nn.GroupNorm(norm_num_groups, channels[0], eps=norm_eps, affine=True),
            nn.SiLU(),
            zero_module(ConvOnly(channels[0], out_channels, 3, 1, 1)),
        )

    def forward(self, x, timesteps, context=None, class_labels=None,
                down_block_additional_residuals=None, mid_block_additional_residual=None):
        if context is not None and not self.with_conditioning:
            raise ValueError("model should have with_conditioning = True if context is provided")
        t_emb = get_timestep_embedding(timesteps, self.block_out_channels[0]).to(dtype=x.dtype)
        emb = self.time_embed(t_emb)
        if self.num_class_embeds is not None:
            if class_labels is None:
                raise ValueError("class_labels should be provided when num_class_embeds > 0")
            emb = emb + self.class_embedding(class_labels).to(dtype=x.dtype)
        h = self.conv_in(x)
        skips = [h]
        for blk in self.down_blocks:
            h, outs = blk(h, emb, context)
            skips.extend(outs)
        if down_block_additional_residuals is not None:
            skips = [s + r for s, r in zip(skips, down_block_additional_residuals)]
        h = self.middle_block(h, emb, context)
        if mid_block_additional_residual is not None:
            h = h + mid_block_additional_residual
        for blk in self.up_blocks:
            h = blk(h, skips, emb, context)
        return self.out(h)


# --------------------------------------------------------------------------------------
# AutoencoderKL  
# --------------------------------------------------------------------------------------
class AEResBlock(nn.Module):
    def __init__(self, cin, cout, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps, affine=True)
        self.conv1 = ConvOnly(cin, cout, 3, 1, 1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps, affine=True)
        self.conv2 = ConvOnly(cout, cout, 3, 1, 1)
        self.nin_shortcut = nn.Identity() if cin == cout else ConvOnly(cin, cout, 1, 1, 0)

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return self.nin_shortcut(x) + h


class AEDownsample(nn.Module):
    """Asymmetric zero pad (end side of each spatial dim) then conv3 stride 2 pad 0."""

    def __init__(self, ch):
        super().__init__()
        self.conv = ConvOnly(ch, ch, 3, 2, 0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1, 0, 1), mode="constant", value=0.0))


class Encoder(nn.Module):
    def __init__(self, in_channels, channels, out_channels, num_res_blocks, groups, eps,
                 attention_levels, with_nonlocal_attn):
        super().__init__()
        blocks: list[nn.Module] = [ConvOnly(in_channels, channels[0], 3, 1, 1)]
        cout = channels[0]
        for i, ch in enumerate(channels):
            cin, cout = cout, ch
            for _ in range(num_res_blocks[i]):
                blocks.append(AEResBlock(cin, cout, groups, eps))
                cin = cout
                if attention_levels[i]:
                    blocks.append(SpatialAttentionBlock(cin, None, groups, eps))
            if i != len(channels) - 1:
                blocks.append(AEDownsample(cin))
        if with_nonlocal_attn:
            blocks += [AEResBlock(channels[-1], channels[-1], groups, eps),
                       SpatialAttentionBlock(channels[-1], None, groups, eps),
                       AEResBlock(channels[-1], channels[-1], groups, eps)]
        blocks.append(nn.GroupNorm(groups, channels[-1], eps=eps, affine=True))
        blocks.append(ConvOnly(channels[-1], out_channels, 3, 1, 1))
        self.blocks = nn.ModuleList(blocks)

    def forward(self, x):
        for b in self.blocks:
            x = b(x)
        return x


class Decoder(nn.Module):
    def __init__(self, channels, in_channels, out_channels, num_res_blocks, groups, eps,
                 attention_levels, with_nonlocal_attn):
        super().__init__()
        rc = list(reversed(channels))
        ra, rn = list(reversed(attention_levels)), list(reversed(num_res_blocks))
        blocks: list[nn.Module] = [ConvOnly(in_channels, rc[0], 3, 1, 1)]
        if with_nonlocal_attn:
            blocks += [AEResBlock(rc[0], rc[0], groups, eps),
                       SpatialAttentionBlock(rc[0], None, groups, eps),
                       AEResBlock(rc[0], rc[0], groups, eps)]
        cout = rc[0]
        for i, ch in enumerate(rc):
            cin, cout = cout, ch
            for _ in range(rn[i]):
                blocks.append(AEResBlock(cin, cout, groups, eps))
                cin = cout
                if ra[i]:
                    blocks.append(SpatialAttentionBlock(cin, None, groups, eps))
            if i != len(rc) - 1:
                blocks.append(_NearestUp(cin))
        blocks.append(nn.GroupNorm(groups, cin, eps=eps, affine=True))
        blocks.append(ConvOnly(cin, out_channels, 3, 1, 1))
        self.blocks = nn.ModuleList(blocks)

    def forward(self, x):
        for b in self.blocks:
            x = b(x)
        return x


class AutoencoderKL(nn.Module):
    def __init__(
        self,
        spatial_dims: int,
        in_channels: int = 1,
        out_channels: int = 1,
        num_res_blocks: Sequence[int] | int = (2, 2, 2, 2),
        channels: Sequence[int] = (32, 64, 64, 64),
        attention_levels: Sequence[bool] = (False, False, True, True),
        latent_channels: int = 3,
        norm_num_groups: int = 32,
        norm_eps: float = 1e-6,
        with_encoder_nonlocal_attn: bool = True,
        with_decoder_nonlocal_attn: bool = True,
        use_checkpoint: bool = False,
        use_convtranspose: bool = False,
        include_fc: bool = True,
        use_combined_linear: bool = False,
        use_flash_attention: bool = False,
    ):
        super().__init__()
        if spatial_dims != 3:
            raise ValueError("oracle restates the 3-D path only")
        if use_convtranspose:
            raise NotImplementedError("use_convtranspose is not used by any reference config")
        if any(c % norm_num_groups for c in channels):
            raise ValueError("AutoencoderKL expects all channels being multiple of norm_num_groups")
        if len(channels) != len(attention_levels):
            raise ValueError("AutoencoderKL expects channels being same size of attention_levels")
        nres = _per_level(num_res_blocks, len(channels))
        if len(nres) != len(channels):
            raise ValueError("`num_res_blocks` should be a single integer or a tuple of integers with the same length as `channels`.")
        self.encoder = Encoder(in_channels, channels, latent_channels, nres, norm_num_groups, norm_eps,
                               attention_levels, with_encoder_nonlocal_attn)
        self.decoder = Decoder(channels, latent_channels, out_channels, nres, norm_num_groups, norm_eps,
                               attention_levels, with_decoder_nonlocal_attn)
        self.quant_conv_mu = ConvOnly(latent_channels, latent_channels, 1, 1, 0)
        self.quant_conv_log_sigma = ConvOnly(latent_channels, latent_channels, 1, 1, 0)
        self.post_quant_conv = ConvOnly(latent_channels, latent_channels, 1, 1, 0)
        self.latent_channels = latent_channels

    def encode(self, x):
        h = self.encoder(x)
        z_mu = self.quant_conv_mu(h)
        z_log_var = torch.clamp(self.quant_conv_log_sigma(h), -30.0, 20.0)
        return z_mu, torch.exp(z_log_var / 2)

    def sampling(self, z_mu, z_sigma):
        return z_mu + torch.randn_like(z_sigma) * z_sigma

    def decode(self, z):
        return self.decoder(self.post_quant_conv(z))

    def reconstruct(self, x):
        return self.decode(self.encode(x)[0])

    def forward(self, x):
        z_mu, z_sigma = self.encode(x)
        return self.decode(self.sampling(z_mu, z_sigma)), z_mu, z_sigma

    def encode_stage_2_inputs(self, x):
        return self.sampling(*self.encode(x))

    def decode_stage_2_outputs(self, z):
        return self.decode(z)


# --------------------------------------------------------------------------------------
# seeded non-zero re-initialisation (SURVEY.md 8d "Synthetic inputs")
# --------------------------------------------------------------------------------------
@torch.no_grad()
def randomize_parameters(model: nn.Module, seed: int = 1234) -> nn.Module:
    """Re-draw EVERY parameter so the zero-initialised convs (UNet conv2 / out.2) stop
    predicting exactly 0: conv/linear W ~ N(0, 1/fan_in), b ~ N(0, 0.02^2),
    GroupNorm gamma ~ 1 + N(0, 0.1^2), beta ~ N(0, 0.1^2).  Deterministic in `seed`
    and in the module traversal order (named_modules)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    for _, m in model.named_modules():
        if isinstance(m, (nn.Conv3d, nn.Linear)):
            fan_in = m.weight[0].numel()
            m.weight.copy_(torch.randn(m.weight.shape, generator=g) / math.sqrt(fan_in))
            if m.bias is not None:
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.02)
        elif isinstance(m, (nn.GroupNorm, nn.LayerNorm)):
            m.weight.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
            m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
        elif isinstance(m, nn.Embedding):
            m.weight.copy_(torch.randn(m.weight.shape, generator=g))
    return model
