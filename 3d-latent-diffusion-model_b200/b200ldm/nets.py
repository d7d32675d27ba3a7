"""B200 drop-ins for the two MONAI networks the reference instantiates by `_target_`
(3d_ldm/config/config_train_32g.json:8,41 through 3d_ldm/utils.py:243-246):

    DiffusionModelUNet(...).forward(x, timesteps, context=None, ...)        3d_ldm/inference.py:75
    AutoencoderKL(...).encode / sampling / decode / decode_stage_2_outputs  3d_ldm/inference.py:71

Same constructor kwargs, forward signatures and state_dict key names as MONAI >= 1.4
(SURVEY.md App. A, B, E), so `load_state_dict(torch.load("diffusion_unet.pt"))`
(3d_ldm/inference.py:72-77) works unchanged.  Parameters are kept in PyTorch layout (fp32);
on first use they are repacked once per dtype into the kernels' [tap][Cout][Cin] layout.
All arithmetic runs in libb200ldm (ops.py); activations are NDHWC internally and NCDHW fp32
at the module boundary.  `precision` = "bf16" (tcgen05 path) or "fp32" (validation path).
"""
from __future__ import annotations

import os
from typing import Sequence

import torch
import torch.nn as nn

from . import ops

_PREC = {"bf16": torch.bfloat16, "fp32": torch.float32}
# GroupNorm partial sums accumulated in the producing conv's epilogue.  Measured (profiles/r01_ncu_conv3d_tc_first.md):
# +0.35 ms per big conv vs 0.06 ms for the standalone stats kernel while the epilogue is not overlapped with the
# MMA main loop, so it is opt-in for now.
#   Superseded by the per-8-channel "chunk stats" of the TMA epilogue (ops.FUSE_GN_STATS, on by default); this flag
#   only keeps the older per-group variant of the direct-store epilogue reachable for experiments.
FUSE_GN_STATS = os.environ.get("B200LDM_FUSE_GN_GROUPSTATS", "0") != "0"


FUSE_UPSAMPLE = os.environ.get("B200LDM_FUSE_UPSAMPLE", "1") != "0"
# narrow-output 3x3x3 convs (Cout <= 16) as a 1x1x1 tensor-core GEMM over stacked taps + a gather (ops.tap_gather):
# the direct kernel re-reads its input 27 times for 16 output columns (0.58 ms for the UNet's out conv, 135 TFLOP/s)
TAP_EXPAND = os.environ.get("B200LDM_TAP_EXPAND", "1") != "0"
# narrow-INPUT 3x3x3 convs (Cin <= 2: the AutoencoderKL's conv_in) as ops.pack_taps + a 1x1x1 GEMM: one k-iteration instead
# of 27 over a zero-padded 64-channel input (0.93 -> ~0.3 ms at 160^3)
TAP_PACK = os.environ.get("B200LDM_TAP_PACK", "1") != "0"
# DiffusionModelUNet.forward (the public module call) replays a cached CUDA graph from the second call of a signature on
FORWARD_GRAPH = os.environ.get("B200LDM_FORWARD_GRAPH", "1") != "0"
# sub-batch streams of the top UNet level inside the captured forward graph (see forward_ndhwc(split=...))
FORWARD_STREAMS = int(os.environ.get("B200LDM_BATCH_STREAMS", "2"))
# ResNet 1x1x1 skip_connection / nin_shortcut accumulated inside conv2 (extra K iterations) on the tensor-core path
FUSE_SKIP = os.environ.get("B200LDM_FUSE_SKIP", "1") != "0"


def default_precision() -> str:
    return os.environ.get("B200LDM_PRECISION", "bf16")


def _pad64(c: int) -> int:
    return ops.round_up(c, 64)


class PackedCache:
    """Per-module cache of repacked weights keyed by (dtype, device, parameter versions)."""

    def __init__(self):
        self._store = {}

    def get(self, key, params, build):
        ver = tuple((p._version, p.data_ptr()) for p in params)
        hit = self._store.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        val = build()
        self._store[key] = (ver, val)
        return val


def param_signature(module: nn.Module) -> int:
    """Changes whenever any parameter is rewritten in place (load_state_dict, optimizer step) or moved: captured
    CUDA graphs bake in the packed-weight pointers, so they are keyed by this."""
    return hash(tuple((p.data_ptr(), p._version) for p in module.parameters()))


class Conv(nn.Module):
    """Leaf that owns `weight` [Cout,Cin,k,k,k] + `bias` like nn.Conv3d and runs the conv3d kernel."""

    def __init__(self, cin: int, cout: int, ksize: int = 3, stride: int = 1, pad_lo: int = 1, pad_hi: int | None = None,
                 terminal: bool = False):
        super().__init__()
        self.cin, self.cout, self.ksize, self.stride = cin, cout, ksize, stride
        self.pad_lo, self.pad_hi = pad_lo, pad_lo if pad_hi is None else pad_hi
        self.weight = nn.Parameter(torch.empty(cout, cin, ksize, ksize, ksize))
        self.bias = nn.Parameter(torch.empty(cout))
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        bound = 1.0 / (cin * ksize ** 3) ** 0.5
        nn.init.uniform_(self.bias, -bound, bound)
        # physical (padded) output channels: feeds another conv -> multiple of 64, else multiple of 16
        self.cout_phys = ops.round_up(cout, 16) if terminal else _pad64(cout)
        self.terminal = terminal
        self._cache = PackedCache()
        # groups of the GroupNorm that consumes this conv's output (set by the owning network): with fewer than 8
        # channels per group (AutoencoderKL: 64 / 128 channels in 32 groups) the epilogue's statistics are emitted
        # per group instead of per 8 channels, so that GroupNorm still needs no statistics pass
        self.stats_groups = None

    def stats_width(self) -> int:
        g = self.stats_groups
        if g and self.cout % g == 0 and self.cout // g in (2, 4) and self.cout_phys % 64 == 0:
            return self.cout // g
        return 8

    def packed(self, dtype, c1: int | None = None, c1_phys: int | None = None, cin_phys: int | None = None):
        """(weights [k^3][cout_phys][cin_phys], bias fp32 [cout_phys]); c1/c1_phys describe a concat split."""
        if cin_phys is None:
            cin_phys = _pad64(self.cin) if c1 is None else c1_phys + _pad64(self.cin - c1)

        def build():
            w = ops.repack_conv_weight(self.weight, self.cout_phys, cin_phys, dtype, c1, c1_phys)
            b = torch.zeros(self.cout_phys, dtype=torch.float32, device=self.weight.device)
            b[: self.cout] = self.bias.detach().float()
            return w, b

        return self._cache.get((dtype, c1, c1_phys, cin_phys), (self.weight, self.bias), build)

    def run(self, x, x2=None, c1: int | None = None, temb=None, temb_stride: int = 0, residual=None, gn_stats=None,
            gn_groups: int = 0, out_f32: bool = False, want_stats: bool = False, sched=None, out=None, cstats_out=None,
            defer: bool = False):
        if self.tap_expanded(x) and x2 is None and temb is None and residual is None and gn_stats is None and not out_f32:
            return self._run_tap_expanded(x, sched, out)
        if sched is not None:
            raise RuntimeError("the fused scheduler update needs the stacked-tap output conv (Conv.tap_expanded)")
        if self.tap_packed(x) and x2 is None and gn_stats is None and not out_f32:
            # 1 or 2 input channels: the 27 taps become the channels of ONE k-iteration (ops.pack_taps) instead of 27
            # k-iterations over a zero-padded 64-channel input
            def build():
                wm = self.weight.detach().float().permute(0, 2, 3, 4, 1).reshape(self.cout, 27 * self.cin)  # [co][t*cin+c]
                wk = ops.repack_conv_weight(wm.contiguous(), self.cout_phys, 64, x.dtype)
                bb = torch.zeros(self.cout_phys, dtype=torch.float32, device=wm.device)
                bb[: self.cout] = self.bias.detach().float()
                return wk, bb

            wk, bb = self._cache.get(("tappack", x.dtype), (self.weight, self.bias), build)
            return ops.conv3d(ops.pack_taps(x, self.cin), wk, bb, self.cout_phys, 1, 1, 0, 0, temb=temb,
                              temb_stride=temb_stride, residual=residual, want_stats=want_stats, out=out,
                              cstats_out=cstats_out, stats_width=self.stats_width(), defer=defer)
        if x2 is None:
            w, b = self.packed(x.dtype, cin_phys=x.shape[-1])
        else:
            w, b = self.packed(x.dtype, c1=c1, c1_phys=x.shape[-1], cin_phys=x.shape[-1] + x2.shape[-1])
        return ops.conv3d(x, w, b, self.cout_phys, self.ksize, self.stride, self.pad_lo, self.pad_hi, x2=x2, temb=temb,
                          temb_stride=temb_stride, residual=residual, gn_stats=gn_stats, gn_groups=gn_groups,
                          out_f32=out_f32, want_stats=want_stats, out=out, cstats_out=cstats_out,
                          stats_width=self.stats_width(), defer=defer)


    def tap_packed(self, x) -> bool:
        """True when this conv runs as ops.pack_taps + a 1x1x1 GEMM for activations like `x` (AutoencoderKL conv_in)"""
        return (TAP_PACK and self.cin <= 2 and self.ksize == 3 and self.stride == 1 and self.pad_lo == 1
                and self.pad_hi == 1 and x.dtype == torch.bfloat16)

    def tap_expanded(self, x) -> bool:
        """True when this conv runs as a 1x1x1 GEMM over stacked taps + ops.tap_gather for activations like `x`."""
        return (TAP_EXPAND and self.terminal and self.cout <= 16 and self.ksize == 3 and self.stride == 1
                and self.pad_lo == 1 and self.pad_hi == 1 and x.dtype == torch.bfloat16)

    def _run_tap_expanded(self, x, sched=None, out=None):
        """3x3x3 conv with <= 16 output channels = 1x1x1 GEMM to 27*cp stacked tap outputs + gather (ops.tap_gather);
        sched: the scheduler update is applied by the gather kernel and the conv output is not materialised."""
        cin_phys = x.shape[-1]
        cp = 1
        while cp < self.cout:
            cp *= 2
        nz = _pad64(27 * cp)

        def build():
            w = self.weight.detach().float()                                   # [Cout, Cin, 3, 3, 3]
            wt = torch.zeros((27, cp, self.cin), dtype=torch.float32, device=w.device)
            wt[:, : self.cout] = w.permute(2, 3, 4, 0, 1).reshape(27, self.cout, self.cin)
            wz = ops.repack_conv_weight(wt.reshape(27 * cp, self.cin), nz, cin_phys, x.dtype)
            b = torch.zeros(16, dtype=torch.float32, device=w.device)
            b[: self.cout] = self.bias.detach().float()
            return wz, b

        wz, b = self._cache.get(("tapexp", x.dtype, cin_phys), (self.weight, self.bias), build)
        z = ops.conv3d(x, wz, None, nz, ksize=1, stride=1, pad_lo=0)
        return ops.tap_gather(z, b, cp, self.cout_phys, sched=sched, want_y=sched is None, out=out)

    def run_with_skip(self, x, skip_conv: "Conv", sx, sx2=None, want_stats: bool = False, defer: bool = False):
        """self(x) + skip_conv(cat(sx, sx2)) in ONE kernel: the 1x1x1 skip conv is a few more K iterations into the same
        TMEM accumulator (no second launch, no residual round trip)."""
        w, _ = self.packed(x.dtype, cin_phys=x.shape[-1])
        c1 = None if sx2 is None else sx.shape[-1]
        cs = sx.shape[-1] + (0 if sx2 is None else sx2.shape[-1])
        sw, _ = skip_conv.packed(x.dtype, c1=c1, c1_phys=c1, cin_phys=cs)

        def build():
            b = torch.zeros(self.cout_phys, dtype=torch.float32, device=self.weight.device)
            b[: self.cout] = self.bias.detach().float() + skip_conv.bias.detach().float()
            return b

        b = self._cache.get(("bias+skip",), (self.bias, skip_conv.bias), build)
        return ops.conv3d(x, w, b, self.cout_phys, self.ksize, self.stride, self.pad_lo, self.pad_hi,
                          skip=(sx, sx2, sw), want_stats=want_stats, stats_width=self.stats_width(), defer=defer)

    def run_upsampled(self, x):
        """conv(nearest_upsample_x2(x)) without materialising the upsampled tensor: on the tensor-core
        path the 27 taps collapse to 8 per output phase (3.4x fewer MACs, weights pre-summed per phase)."""
        cin_phys = x.shape[-1]
        if x.dtype == torch.bfloat16 and FUSE_UPSAMPLE:
            def build():
                w = ops.repack_upconv_weight(self.weight, self.cout_phys, cin_phys, x.dtype)
                b = torch.zeros(self.cout_phys, dtype=torch.float32, device=self.weight.device)
                b[: self.cout] = self.bias.detach().float()
                return w, b

            w, b = self._cache.get(("up", x.dtype, cin_phys), (self.weight, self.bias), build)
        else:
            w, b = self.packed(x.dtype, cin_phys=cin_phys)
            if not FUSE_UPSAMPLE:
                return ops.conv3d(ops.upsample_nearest2x(x), w, b, self.cout_phys, 3, 1, 1, 1, want_stats=True,
                                  stats_width=self.stats_width())
        return ops.conv3d(x, w, b, self.cout_phys, 3, 1, 1, 1, upsample2x=True, want_stats=True,
                          stats_width=self.stats_width())


class ConvOnly(nn.Module):
    """MONAI `Convolution(conv_only=True)`: container with a single child `.conv` (key `*.conv.weight`)."""

    def __init__(self, cin, cout, ksize=3, stride=1, pad_lo=1, pad_hi=None, terminal=False):
        super().__init__()
        self.conv = Conv(cin, cout, ksize, stride, pad_lo, pad_hi, terminal)

    def run(self, *a, **k):
        return self.conv.run(*a, **k)


class Linear(nn.Module):
    """nn.Linear-shaped parameters; runs as a 1x1x1 conv (GEMM) in fp32 or bf16."""

    def __init__(self, cin: int, cout: int, bias: bool = True):
        super().__init__()
        self.cin, self.cout = cin, cout
        self.weight = nn.Parameter(torch.empty(cout, cin))
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        if bias:
            self.bias = nn.Parameter(torch.empty(cout))
            nn.init.uniform_(self.bias, -1.0 / cin ** 0.5, 1.0 / cin ** 0.5)
        else:
            self.register_parameter("bias", None)
        self._cache = PackedCache()

    def packed(self, dtype, cin_phys: int | None = None):
        """(weights [1][round16(cout)][cin_phys], bias fp32 | None); cin_phys > cin: zero columns for a padded input."""
        cin_phys = cin_phys or self.cin

        def build():
            w = ops.repack_conv_weight(self.weight, ops.round_up(self.cout, 16), cin_phys, dtype)
            if self.bias is None:
                return w, None
            b = torch.zeros(ops.round_up(self.cout, 16), dtype=torch.float32, device=self.weight.device)
            b[: self.cout] = self.bias.detach().float()
            return w, b

        ps = (self.weight,) if self.bias is None else (self.weight, self.bias)
        return self._cache.get((dtype, cin_phys), ps, build)


class GroupNorm(nn.Module):
    def __init__(self, groups: int, channels: int, eps: float):
        super().__init__()
        self.groups, self.channels, self.eps = groups, channels, eps
        self.weight = nn.Parameter(torch.ones(channels))
        self.bias = nn.Parameter(torch.zeros(channels))
        self._cache = PackedCache()

    def run(self, x, silu: bool, x2=None, stats=None):
        cphys = x.shape[-1] + (0 if x2 is None else x2.shape[-1])
        if cphys == self.channels:
            return ops.groupnorm(x, self.weight.detach(), self.bias.detach(), self.groups, self.eps, silu, x2=x2,
                                 stats=stats)
        # channel-padded activation (a 32-channel level lives in a 64-channel tensor whose upper half is exactly zero,
        # config_train_stable.json:13-27): the kernel sees cphys / (channels per group) groups - the real ones followed by
        # all-zero padding groups - with gamma = beta = 0 on the padding, so padded channels stay exactly zero
        cpg = self.channels // self.groups
        if x2 is not None or stats is not None or cphys % cpg:
            raise NotImplementedError("GroupNorm over channel-padded tensors: single source, no pre-reduced statistics")

        def build():
            g = torch.zeros(cphys, dtype=torch.float32, device=self.weight.device)
            b = torch.zeros(cphys, dtype=torch.float32, device=self.weight.device)
            g[: self.channels] = self.weight.detach().float()
            b[: self.channels] = self.bias.detach().float()
            return g, b

        g, b = self._cache.get(("padded", cphys), (self.weight, self.bias), build)
        return ops.groupnorm(x, g, b, cphys // cpg, self.eps, silu)


def zero_module(m: nn.Module) -> nn.Module:
    for p in m.parameters():
        nn.init.zeros_(p)
    return m


# ------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------
class SABlock(nn.Module):
    """Separate to_q/to_k/to_v/out_proj linears (MONAI key names); executed as one fused QKV GEMM."""

    def __init__(self, hidden: int, heads: int):
        super().__init__()
        if hidden % heads:
            raise ValueError("hidden size should be divisible by num_heads.")
        self.hidden, self.heads, self.head_dim = hidden, heads, hidden // heads
        self.scale = self.head_dim ** -0.5
        self.to_q, self.to_k, self.to_v = Linear(hidden, hidden), Linear(hidden, hidden), Linear(hidden, hidden)
        self.out_proj = Linear(hidden, hidden)
        self._cache = PackedCache()

    def packed_qkv(self, dtype):
        ps = [self.to_q.weight, self.to_k.weight, self.to_v.weight, self.to_q.bias, self.to_k.bias, self.to_v.bias]

        def build():
            w = torch.cat([self.to_q.weight, self.to_k.weight, self.to_v.weight], dim=0).detach()
            b = torch.cat([self.to_q.bias, self.to_k.bias, self.to_v.bias], dim=0).detach().float().contiguous()
            return ops.repack_conv_weight(w, 3 * self.hidden, self.hidden, dtype), b

        return self._cache.get(dtype, ps, build)

    def run(self, xn, residual):
        """xn: normalised tokens [B, D,H,W, C]; residual: same shape; returns out_proj(att) + residual."""
        shp = xn.shape
        B, Cc = shp[0], shp[-1]
        L = xn.numel() // (B * Cc)
        wqkv, bqkv = self.packed_qkv(xn.dtype)
        qkv = ops.linear(xn.reshape(B * L, Cc), wqkv, bqkv, 3 * Cc)
        att = ops.attention(qkv.reshape(B, L, 3 * Cc), self.heads, self.head_dim, self.scale)
        wo, bo = self.out_proj.packed(xn.dtype)
        # 1x1x1 conv on the 5-D view (keeps the per-sample axis so the epilogue can emit GroupNorm chunk stats)
        # defer: the block output goes straight into the next block's GroupNorm (see ops.PendingFinalize)
        return ops.conv3d(att.reshape(shp), wo, bo, Cc, ksize=1, stride=1, pad_lo=0, residual=residual, want_stats=True,
                          defer=True)


class SpatialAttentionBlock(nn.Module):
    def __init__(self, channels: int, num_head_channels: int | None, groups: int, eps: float):
        super().__init__()
        self.norm = GroupNorm(groups, channels, eps)
        heads = channels // num_head_channels if num_head_channels else 1
        self.attn = SABlock(channels, heads)

    def run(self, x):
        return self.attn.run(self.norm.run(x, silu=False), x)


class LayerNorm(nn.Module):
    def __init__(self, channels: int, eps: float = 1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(channels))
        self.bias = nn.Parameter(torch.zeros(channels))

    def run(self, x2d):
        return ops.layernorm(x2d, self.weight.detach(), self.bias.detach(), self.eps)


class XAttention(nn.Module):
    """attn1 / attn2 of MONAI DiffusionUNetTransformerBlock (SABlock / CrossAttentionBlock with qkv_bias=False):
    q from the tokens, k/v from the context (the tokens themselves for attn1).  attn1 runs one fused QKV GEMM,
    attn2 one Q GEMM + one fused KV GEMM over the context; out_proj carries the residual in its epilogue."""

    def __init__(self, hidden: int, heads: int, head_dim: int, context_dim: int | None = None):
        super().__init__()
        inner = heads * head_dim
        self.hidden, self.inner, self.heads, self.head_dim = hidden, inner, heads, head_dim
        self.scale = head_dim ** -0.5
        self.context_dim = context_dim
        self.to_q = Linear(hidden, inner, bias=False)
        self.to_k = Linear(context_dim or hidden, inner, bias=False)
        self.to_v = Linear(context_dim or hidden, inner, bias=False)
        self.out_proj = Linear(inner, hidden)
        self._cache = PackedCache()

    def _packed_cat(self, key, linears, dtype, cin_phys):
        ps = [l.weight for l in linears]

        def build():
            w = torch.cat(ps, dim=0).detach()
            return ops.repack_conv_weight(w, w.shape[0], cin_phys, dtype)

        return self._cache.get((key, dtype, cin_phys), ps, build)

    def run(self, xn, residual, B: int, L: int, ctx=None):
        """xn, residual: [B*L, hidden]; ctx: None (self attention) or [B*Lc, pad64(context_dim)] context tokens."""
        dt = xn.dtype
        if ctx is None:
            wqkv = self._packed_cat("qkv", (self.to_q, self.to_k, self.to_v), dt, self.hidden)
            qkv = ops.linear(xn, wqkv, None, 3 * self.inner)
            att = ops.attention(qkv.reshape(B, L, 3 * self.inner), self.heads, self.head_dim, self.scale)
        else:
            wq, _ = self.to_q.packed(dt)
            wkv = self._packed_cat("kv", (self.to_k, self.to_v), dt, ctx.shape[-1])
            q = ops.linear(xn, wq, None, self.inner)
            kv = ops.linear(ctx, wkv, None, 2 * self.inner)
            att = ops.cross_attention(q.reshape(B, L, self.inner), kv.reshape(B, ctx.shape[0] // B, 2 * self.inner),
                                      self.heads, self.head_dim, self.scale)
        wo, bo = self.out_proj.packed(dt)
        return ops.linear(att.reshape(B * L, self.inner), wo, bo, self.hidden, residual=residual)


class GEGLUMLP(nn.Module):
    """MONAI MLPBlock(act="GEGLU"): linear1 -> value * gelu(gate) -> linear2 (+ residual in the GEMM epilogue)."""

    def __init__(self, hidden: int, mlp_dim: int):
        super().__init__()
        self.linear1 = Linear(hidden, mlp_dim * 2)
        self.linear2 = Linear(mlp_dim, hidden)

    def run(self, xn, residual):
        w1, b1 = self.linear1.packed(xn.dtype)
        w2, b2 = self.linear2.packed(xn.dtype)
        h = ops.geglu(ops.linear(xn, w1, b1, self.linear1.cout))
        return ops.linear(h, w2, b2, self.linear2.cout, residual=residual)


class TransformerBlock(nn.Module):
    def __init__(self, channels: int, heads: int, head_dim: int, cross_attention_dim: int | None):
        super().__init__()
        self.attn1 = XAttention(channels, heads, head_dim)
        self.ff = GEGLUMLP(channels, channels * 4)
        self.attn2 = XAttention(channels, heads, head_dim, cross_attention_dim)
        self.norm1, self.norm2, self.norm3 = LayerNorm(channels), LayerNorm(channels), LayerNorm(channels)

    def run(self, t, B: int, L: int, ctx):
        t = self.attn1.run(self.norm1.run(t), t, B, L)
        t = self.attn2.run(self.norm2.run(t), t, B, L, ctx)
        return self.ff.run(self.norm3.run(t), t)


class SpatialTransformer(nn.Module):
    """MONAI SpatialTransformer (attention levels of a with_conditioning=True UNet): GroupNorm -> 1x1x1 proj_in ->
    transformer blocks over the D*H*W tokens (NDHWC activations ARE the token matrix, no permute) -> 1x1x1 proj_out
    (zero-init) + residual."""

    def __init__(self, channels: int, heads: int, head_dim: int, num_layers: int, groups: int, eps: float,
                 cross_attention_dim: int | None):
        super().__init__()
        inner = heads * head_dim
        if inner % 64:
            raise NotImplementedError("b200ldm SpatialTransformer needs heads * head_dim to be a multiple of 64")
        self.norm = GroupNorm(groups, channels, eps)
        self.proj_in = ConvOnly(channels, inner, 1, 1, 0)
        self.transformer_blocks = nn.ModuleList(
            [TransformerBlock(inner, heads, head_dim, cross_attention_dim) for _ in range(num_layers)])
        self.proj_out = zero_module(ConvOnly(inner, channels, 1, 1, 0))

    def run(self, x, ctx):
        shp = x.shape
        B = shp[0]
        t = self.proj_in.run(self.norm.run(x, silu=False))
        inner = t.shape[-1]
        L = t.numel() // (B * inner)
        t2 = t.reshape(B * L, inner)
        for blk in self.transformer_blocks:
            t2 = blk.run(t2, B, L, ctx)
        return self.proj_out.run(t2.reshape(shp[:-1] + (inner,)), residual=x, want_stats=True)


def _make_attn(cout, head_ch, groups, eps, xattn):
    """xattn None: SpatialAttentionBlock; (num_layers, cross_attention_dim): SpatialTransformer."""
    if xattn is None:
        return SpatialAttentionBlock(cout, head_ch, groups, eps)
    return SpatialTransformer(cout, cout // head_ch, head_ch, xattn[0], groups, eps, xattn[1])


class NearestUp(nn.Module):
    """UpSample(mode="nontrainable", nearest x2) + post conv; children `upsample_non_trainable`, `postconv`."""

    def __init__(self, ch: int):
        super().__init__()
        self.upsample_non_trainable = nn.Identity()
        self.postconv = ConvOnly(ch, ch, 3, 1, 1)

    def run(self, x):
        return self.postconv.conv.run_upsampled(x)


# ------------------------------------------------------------------------------------------
# DiffusionModelUNet
# ------------------------------------------------------------------------------------------
class UNetResBlock(nn.Module):
    def __init__(self, cin, cout, temb_ch, groups, eps):
        super().__init__()
        self.cin, self.cout, self.groups = cin, cout, groups
        self.norm1 = GroupNorm(groups, cin, eps)
        self.conv1 = ConvOnly(cin, cout)
        self.time_emb_proj = Linear(temb_ch, cout)
        self.norm2 = GroupNorm(groups, cout, eps)
        self.conv2 = zero_module(ConvOnly(cout, cout))
        self.skip_connection = nn.Identity() if cin == cout else ConvOnly(cin, cout, 1, 1, 0)

    def run(self, x, temb, temb_stride, x2=None):
        """x (and optional skip x2, concatenated on channels); temb: fp32 view [n_t, cout] for this block."""
        h = self.norm1.run(x, silu=True, x2=x2)
        stats = None
        cpg = self.cout // self.groups
        if FUSE_GN_STATS and cpg >= 2 and (cpg & (cpg - 1)) == 0:
            stats = ops.zeros_f64(x.shape[0] * self.groups * 2, x.device)
        # defer: a split-K conv (the 4^3 / 5^3 levels) leaves its finalize pass to the GroupNorm that consumes the tensor -
        # norm2 right here for conv1, the next block's first GroupNorm for the block output (ops.PendingFinalize)
        h = self.conv1.run(h, temb=temb, temb_stride=temb_stride, gn_stats=stats, gn_groups=self.groups,
                           want_stats=stats is None, defer=stats is None)
        h = self.norm2.run(h, silu=True, stats=stats)
        if isinstance(self.skip_connection, nn.Identity):
            return self.conv2.run(h, residual=x, want_stats=True, defer=True)
        if FUSE_SKIP and h.dtype == torch.bfloat16:
            return self.conv2.conv.run_with_skip(h, self.skip_connection.conv, x, x2, want_stats=True, defer=True)
        res = self.skip_connection.run(x, x2=x2, c1=None if x2 is None else x.shape[-1])
        return self.conv2.run(h, residual=res, want_stats=True, defer=True)


class _DownOp(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.op = ConvOnly(ch, ch, 3, 2, 1)


def _alias_upsampler_keys(state_dict, prefix, *_):
    """load_state_dict pre-hook.  MONAI >= 1.4 names the post-upsample conv of a UNet up block
    `up_blocks.{i}.upsampler.postconv.conv.*` (WrappedUpsample subclasses UpSample - an nn.Sequential with children
    `upsample_non_trainable`, `postconv` - and is assigned to `upsampler` itself); that is what state_dict() emits.
    The nested spelling `up_blocks.{i}.upsampler.upsample.postconv.conv.*` loads as an alias."""
    for k in [k for k in state_dict if k.startswith(prefix) and ".upsampler.upsample." in k]:
        state_dict[k.replace(".upsampler.upsample.", ".upsampler.")] = state_dict.pop(k)


class _DownBlock(nn.Module):
    def __init__(self, cin, cout, temb, n_res, groups, eps, add_down, attn, head_ch, xattn=None):
        super().__init__()
        self.resnets = nn.ModuleList([UNetResBlock(cin if j == 0 else cout, cout, temb, groups, eps)
                                      for j in range(n_res)])
        if attn:
            self.attentions = nn.ModuleList([_make_attn(cout, head_ch, groups, eps, xattn) for _ in range(n_res)])
        self.has_attn = attn
        self.downsampler = _DownOp(cout) if add_down else None


class _MidBlock(nn.Module):
    def __init__(self, ch, temb, groups, eps, head_ch, xattn=None):
        super().__init__()
        self.resnet_1 = UNetResBlock(ch, ch, temb, groups, eps)
        self.attention = _make_attn(ch, head_ch, groups, eps, xattn)
        self.resnet_2 = UNetResBlock(ch, ch, temb, groups, eps)


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
        self.has_attn = attn
        self.upsampler = NearestUp(cout) if add_up else None


def load_old_state_dict(module: nn.Module, old_state_dict: dict, verbose: bool = False) -> None:
    """MONAI's `load_old_state_dict` (both nets have it since the MONAI-Generative merge): ingest a checkpoint written by
    the pre-1.4 `generative` package.  Its attention blocks kept `to_q / to_k / to_v / proj_attn` directly on the block
    (now `attn.to_q ... attn.out_proj`), transformer blocks called their output projection `to_out.0` (now `out_proj`),
    and the post-upsample conv was `conv` (now `postconv`).  Everything else has the same name.  Strict: a key that is
    still missing afterwards raises from load_state_dict."""
    old = dict(old_state_dict)
    new = module.state_dict()
    if all(k in new for k in old):
        module.load_state_dict(old)
        return
    renamed = {}
    for k in new:
        if k in old:
            renamed[k] = old.pop(k)
            continue
        cand = k
        if ".attn.out_proj." in k and "transformer_blocks" not in k:
            cand = k.replace(".attn.out_proj.", ".proj_attn.")
        elif ".attn.to_" in k and "transformer_blocks" not in k:
            cand = k.replace(".attn.to_", ".to_")
        elif ".out_proj." in k and "transformer_blocks" in k:
            cand = k.replace(".out_proj.", ".to_out.0.")
        elif "postconv" in k:
            cand = k.replace("postconv", "conv")
        if cand in old:
            renamed[k] = old.pop(cand)
    if verbose:
        for k in new:
            if k not in renamed:
                print(f"key {k} not found in old state dict")
        for k in old:
            print(f"key {k} of the old state dict is unused")
    module.load_state_dict(renamed, strict=True)


def _per_level(v, n):
    return list(v) if isinstance(v, (list, tuple)) else [v] * n


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
        precision: str | None = None,
    ):
        super().__init__()
        if spatial_dims != 3:
            raise ValueError("b200ldm implements the 3-D hot path only (spatial_dims=3)")
        if with_conditioning and cross_attention_dim is None:
            raise ValueError("DiffusionModelUNet expects dimension of the cross-attention conditioning "
                             "(cross_attention_dim) when using with_conditioning.")
        if cross_attention_dim is not None and not with_conditioning:
            raise ValueError("DiffusionModelUNet expects with_conditioning=True when specifying the cross_attention_dim.")
        if resblock_updown:
            raise NotImplementedError("resblock_updown is not used by any reference config")
        if dropout_cattn > 1.0 or dropout_cattn < 0.0:
            raise ValueError("Dropout cannot be negative or >1.0!")
        if any(c % norm_num_groups for c in channels):
            raise ValueError("DiffusionModelUNet expects all channels being multiple of norm_num_groups")
        if len(channels) != len(attention_levels):
            raise ValueError("DiffusionModelUNet expects channels being same size of attention_levels")
        if any(c % 64 for c in channels):
            raise NotImplementedError("b200ldm UNet needs channel counts that are multiples of 64 (128-byte TMA rows)")
        L = len(channels)
        head, nres = _per_level(num_head_channels, L), _per_level(num_res_blocks, L)
        if len(head) != L:
            raise ValueError("num_head_channels should have the same length as attention_levels.")
        if len(nres) != L:
            raise ValueError("`num_res_blocks` should be a single integer or a tuple of integers with the same length "
                             "as `channels`.")
        self.in_channels, self.out_channels = in_channels, out_channels
        self._fwd_graphs = {}
        self._register_load_state_dict_pre_hook(_alias_upsampler_keys)
        self.block_out_channels = list(channels)
        self.with_conditioning = with_conditioning
        self.cross_attention_dim = cross_attention_dim
        self.precision = precision or default_precision()
        temb = channels[0] * 4
        xattn = (transformer_num_layers, cross_attention_dim) if with_conditioning else None

        self.conv_in = ConvOnly(in_channels, channels[0])
        self.time_embed = nn.Sequential(Linear(channels[0], temb), nn.SiLU(), Linear(temb, temb))
        self.num_class_embeds = num_class_embeds
        if num_class_embeds is not None:
            self.class_embedding = nn.Embedding(num_class_embeds, temb)
        self.down_blocks = nn.ModuleList()
        cout = channels[0]
        for i in range(L):
            cin, cout = cout, channels[i]
            self.down_blocks.append(_DownBlock(cin, cout, temb, nres[i], norm_num_groups, norm_eps, i != L - 1,
                                               bool(attention_levels[i]), head[i], xattn))
        self.middle_block = _MidBlock(channels[-1], temb, norm_num_groups, norm_eps, head[-1], xattn)
        self.up_blocks = nn.ModuleList()
        rc, rn = list(reversed(channels)), list(reversed(nres))
        ra, rh = list(reversed(attention_levels)), list(reversed(head))
        cout = rc[0]
        for i in range(L):
            prev, cout = cout, rc[i]
            cin = rc[min(i + 1, L - 1)]
            self.up_blocks.append(_UpBlock(cin, prev, cout, temb, rn[i] + 1, norm_num_groups, norm_eps, i != L - 1,
                                           bool(ra[i]), rh[i], xattn))
        self.out = nn.Sequential(GroupNorm(norm_num_groups, channels[0], norm_eps), nn.SiLU(),
                                 zero_module(ConvOnly(channels[0], out_channels, terminal=True)))
        self._cache = PackedCache()

    def load_old_state_dict(self, old_state_dict: dict, verbose: bool = False) -> None:
        load_old_state_dict(self, old_state_dict, verbose)

    # ---- time embedding (always fp32: tiny M, and it keeps t-conditioning exact in both modes)
    def resblocks(self):
        out = []
        for b in self.down_blocks:
            out += list(b.resnets)
        out += [self.middle_block.resnet_1, self.middle_block.resnet_2]
        for b in self.up_blocks:
            out += list(b.resnets)
        return out

    def _temb_layout(self):
        offs, o = [], 0
        for r in self.resblocks():
            offs.append(o)
            o += r.cout
        return offs, o

    def time_embedding_table(self, timesteps: torch.Tensor, class_labels: torch.Tensor | None = None) -> torch.Tensor:
        """[n_t] -> fp32 [n_t, sum(C_out over ResNet blocks)]: time_emb_proj_i(SiLU(time_embed(sincos(t)))) for every
        block i, computed by one fused GEMM (the 19 M=1 GEMVs are independent of x).  With
        num_class_embeds the class embedding of row r's label is added to the time embedding first."""
        if self.num_class_embeds is not None and class_labels is None:
            raise ValueError("class_labels should be provided when num_class_embeds > 0")
        rbs = self.resblocks()
        params = [p for r in rbs for p in (r.time_emb_proj.weight, r.time_emb_proj.bias)]

        def build():
            w = torch.cat([r.time_emb_proj.weight for r in rbs], dim=0).detach()
            b = torch.cat([r.time_emb_proj.bias for r in rbs], dim=0).detach().float().contiguous()
            return ops.repack_conv_weight(w, ops.round_up(w.shape[0], 16), w.shape[1], torch.float32), b, w.shape[0]

        wall, ball, total = self._cache.get("temb_proj", params, build)
        t_emb = ops.timestep_embedding(timesteps, self.block_out_channels[0])
        l0, l2 = self.time_embed[0], self.time_embed[2]
        w0, b0 = l0.packed(torch.float32)
        w2, b2 = l2.packed(torch.float32)
        h = ops.linear(t_emb, w0, b0, l0.cout)
        h = ops.linear(ops.silu_f32(h), w2, b2, l2.cout)
        if self.num_class_embeds is not None:
            labels = class_labels.to(device=h.device, dtype=torch.int64).reshape(-1)
            if labels.numel() != h.shape[0]:
                if h.shape[0] != 1:
                    raise ValueError("class_labels must have one entry per timestep row")
                h = h.expand(labels.numel(), -1).contiguous()
            h = ops.embedding_add(h, self.class_embedding.weight.detach().float().contiguous(), labels.contiguous())
        return ops.linear(ops.silu_f32(h), wall, ball, total)

    def context_tokens(self, context: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """[B, Lc, cross_attention_dim] (any float dtype) -> [B*Lc, pad64(dim)] in the compute dtype, zero padded
        (the K/V projections read it through the same TMA path as every other GEMM operand)."""
        if context.ndim != 3 or context.shape[-1] != self.cross_attention_dim:
            raise ValueError(f"context must be [B, L, {self.cross_attention_dim}], got {tuple(context.shape)}")
        B, Lc, Dc = context.shape
        src = context.detach().to(torch.float32).reshape(B * Lc, Dc, 1, 1, 1).contiguous()
        o = None if out is None else out.reshape(B * Lc, 1, 1, 1, -1)
        return ops.to_ndhwc(src, _pad64(Dc), _PREC[self.precision], out=o).reshape(B * Lc, _pad64(Dc))

    # ---- NDHWC core
    def forward_ndhwc(self, x, temb_all, temb_stride: int, sched=None, ctx=None, split=None):
        """x: [N,D,H,W,pad64(in_channels)] (bf16|fp32); temb_all: fp32 [n_t, total] rows (n_t = 1 -> stride 0);
        ctx: context tokens from `context_tokens` (with_conditioning=True models only).
        Returns eps [N,D,H,W,round16(out_channels)].

        split = (cuts, side_streams): the TOP resolution level - conv_in, down block 0 with its down-sampler, the last
        up block and the output conv - runs as len(cuts) sub-batches [lo, hi) on parallel streams (volumes are
        independent); the lower levels run on the whole batch.  At 8 x 20^3 a top-level conv is 256 CTAs = two waves
        for 1.7 waves of work; two half batches on two streams keep the SMs fed through each other's partial wave
        (6 x (GroupNorm + conv): 1.01 -> 0.92 ms, scripts/ab_level0_streams.py), while the small levels would only pay
        for a split (more split-K, weights read twice).  `sched` is then one b200ldm_sched_params per sub-batch (the
        fused update is the only output, returns None) or None (returns the prediction of the whole batch)."""
        if self.with_conditioning and ctx is None:
            raise ValueError("with_conditioning=True: a context tensor is required (cross-attention conditioning)")
        N = x.shape[0]
        offs, _ = self._temb_layout()
        toff = {id(r): o for r, o in zip(self.resblocks(), offs)}
        lc = 0 if ctx is None else ctx.shape[0] // N

        def temb_for(r, lo=0):   # pointer offset only; the kernel strides rows by temb_stride
            return temb_all[lo if temb_stride else 0:, toff[id(r)]:]

        def attend(a, h, lo=0, hi=N):
            return a.run(h, ctx[lo * lc:hi * lc]) if self.with_conditioning else a.run(h)

        def bslice(t, lo, hi):   # batch slice that keeps the per-sample GroupNorm chunk sums attached
            v = t[lo:hi]
            cs = getattr(t, "cstats", None)
            if cs is not None:
                v.cstats = cs.view(N, -1)[lo:hi].reshape(-1)
                v.cstats_width = getattr(t, "cstats_width", 8)
            return v

        blk0, blk_last = self.down_blocks[0], self.up_blocks[-1]

        def head(lo, hi, down_out):
            """conv_in + down block 0 for volumes [lo, hi); the down-sampled tensor lands in down_out[lo:hi]"""
            h = self.conv_in.run(x[lo:hi], want_stats=True)
            sk = [h]
            for j, r in enumerate(blk0.resnets):
                h = r.run(h, temb_for(r, lo), temb_stride)
                if blk0.has_attn:
                    h = attend(blk0.attentions[j], h, lo, hi)
                sk.append(h)
            if blk0.downsampler is not None:
                if down_out is None:
                    h = blk0.downsampler.op.run(h, want_stats=True)
                else:
                    full, cs, emitted = down_out
                    h = blk0.downsampler.op.run(h, want_stats=True, out=full[lo:hi],
                                                cstats_out=cs.view(N, -1)[lo:hi].reshape(-1))
                    emitted.append(getattr(h, "cstats", None) is not None)   # split-K plans emit no chunk sums
            return h, sk

        def body(h, skips):
            """down blocks 1.., middle block, every up block but the last - whole batch"""
            for blk in self.down_blocks[1:]:
                for j, r in enumerate(blk.resnets):
                    h = r.run(h, temb_for(r), temb_stride)
                    if blk.has_attn:
                        h = attend(blk.attentions[j], h)
                    skips.append(h)
                if blk.downsampler is not None:
                    h = blk.downsampler.op.run(h, want_stats=True, defer=True)   # next consumer: a ResNet block's norm1
                    skips.append(h)
            mb = self.middle_block
            h = mb.resnet_1.run(h, temb_for(mb.resnet_1), temb_stride)
            h = attend(mb.attention, h)
            h = mb.resnet_2.run(h, temb_for(mb.resnet_2), temb_stride)
            for blk in self.up_blocks[:-1]:
                for j, r in enumerate(blk.resnets):
                    h = r.run(h, temb_for(r), temb_stride, x2=skips.pop())
                    if blk.has_attn:
                        h = attend(blk.attentions[j], h)
                if blk.upsampler is not None:
                    h = blk.upsampler.run(h)
            return h

        def tail(h, sk, lo, hi, sp, out=None):
            """last up block + output conv for volumes [lo, hi); sk: that sub-batch's top-level skips"""
            for j, r in enumerate(blk_last.resnets):
                h = r.run(h, temb_for(r, lo), temb_stride, x2=sk.pop())
                if blk_last.has_attn:
                    h = attend(blk_last.attentions[j], h, lo, hi)
            if blk_last.upsampler is not None:
                h = blk_last.upsampler.run(h)
            h = self.out[0].run(h, silu=True)
            return self.out[2].run(h, sched=sp, out=out)   # sched: the scheduler update rides in this conv's last kernel

        if split is None or blk0.downsampler is None:
            if isinstance(sched, (list, tuple)):
                raise ValueError("per-sub-batch scheduler parameters need a split (and a UNet with more than one level)")
            h, top = head(0, N, None)
            h = body(h, [h] if blk0.downsampler is not None else [])
            return tail(h, top, 0, N, sched)

        cuts, side = split
        if len(cuts) != len(side) + 1 or (sched is not None and (not isinstance(sched, (list, tuple))
                                                                 or len(sched) != len(cuts))):
            raise ValueError("split = (cuts, side_streams) needs one stream less than, and one sched per, sub-batch")
        cur = torch.cuda.current_stream()
        # down-sampled tensor of the whole batch, written slice by slice by the sub-batches' down-samplers
        ds = blk0.downsampler.op.conv
        osp = tuple(ops.conv_out_extent(v, 3, 2, ds.pad_lo, ds.pad_hi) for v in x.shape[1:4])
        down_full = torch.empty((N,) + osp + (ds.cout_phys,), dtype=x.dtype, device=x.device)
        down_out = (down_full, ops.zeros_f64(N * (ds.cout_phys // 8) * 2, x.device), [])
        tops = [None] * len(cuts)

        def fork(fn):
            for s in side:
                s.wait_stream(cur)
            for i, s in enumerate(side):
                with torch.cuda.stream(s):
                    fn(i + 1)
            fn(0)
            for s in side:
                cur.wait_stream(s)

        def run_head(i):
            _, tops[i] = head(cuts[i][0], cuts[i][1], down_out)
        fork(run_head)
        if all(down_out[2]):
            down_full.cstats = down_out[1]
        h = body(down_full, [down_full])

        eps = None
        if sched is None:   # no fused scheduler update: the sub-batches write batch slices of one prediction tensor
            oc = self.out[2].conv
            eps = torch.empty(tuple(x.shape[:4]) + (oc.cout_phys,), dtype=x.dtype, device=x.device)

        def run_tail(i):
            lo, hi = cuts[i]
            tail(bslice(h, lo, hi), tops[i], lo, hi, None if sched is None else sched[i],
                 None if eps is None else eps[lo:hi])
        fork(run_tail)
        return eps

    def forward(self, x, timesteps, context=None, class_labels=None, down_block_additional_residuals=None,
                mid_block_additional_residual=None):
        if context is not None and not self.with_conditioning:
            raise ValueError("model should have with_conditioning = True if context is provided")
        if down_block_additional_residuals is not None or mid_block_additional_residual is not None:
            raise NotImplementedError("ControlNet residuals are not on the reference's hot path")
        if self.num_class_embeds is not None and class_labels is None:
            raise ValueError("class_labels should be provided when num_class_embeds > 0")
        if self.num_class_embeds is None:
            class_labels = None
        if x.ndim != 5 or x.shape[1] != self.in_channels:
            raise ValueError(f"expected [B,{self.in_channels},D,H,W], got {tuple(x.shape)}")
        if timesteps.ndim != 1:
            raise ValueError("Timesteps should be a 1d-array")
        if timesteps.numel() not in (1, x.shape[0]):
            raise ValueError("timesteps must have 1 or batch elements")
        if context is not None and context.shape[0] != x.shape[0]:
            raise ValueError("context must have one row of tokens per batch element")
        with torch.no_grad():
            if FORWARD_GRAPH and x.is_cuda and not torch.cuda.is_current_stream_capturing():
                return self._forward_graphed(x, timesteps, context, class_labels)
            return self._forward_eager(x, timesteps, context, class_labels)

    def _forward_eager(self, x, timesteps, context=None, class_labels=None):
        dt = _PREC[self.precision]
        xa = ops.to_ndhwc(x, _pad64(self.in_channels), dt)
        if class_labels is not None and timesteps.numel() == 1 and x.shape[0] > 1:
            timesteps = timesteps.expand(x.shape[0])
        temb_all = self.time_embedding_table(timesteps.to(device=x.device, dtype=torch.float32), class_labels)
        stride = 0 if temb_all.shape[0] == 1 else temb_all.shape[1]
        ctx = None if context is None else self.context_tokens(context.to(x.device))
        eps = self.forward_ndhwc(xa, temb_all, stride, ctx=ctx, split=self._forward_split(xa))
        return ops.to_ncdhw(eps, self.out_channels)

    def _forward_split(self, xa):
        """(cuts, side streams) for the top-level sub-batch split of forward_ndhwc, or None: only inside a graph capture
        (parallel branches; in eager mode the launch overhead of a second stream buys nothing) and only where the
        top level fills the chip more than once (same rule as the sampling loop)."""
        n, vox = xa.shape[0], xa.shape[1] * xa.shape[2] * xa.shape[3]
        live = torch.cuda.is_current_stream_capturing() or self.__dict__.get("_warm_split", False)
        if (FORWARD_STREAMS < 2 or n < 2 or n * vox < 2 * 148 * 128 or not live
                or not self.out[2].conv.tap_expanded(xa) or self.down_blocks[0].downsampler is None):
            return None
        k = min(FORWARD_STREAMS, n)
        cuts = [round(i * n / k) for i in range(k + 1)]
        side = self.__dict__.setdefault("_side_streams", {})
        key = (str(xa.device), k)
        if key not in side:
            side[key] = [torch.cuda.Stream(device=xa.device) for _ in range(k - 1)]
        return [(cuts[i], cuts[i + 1]) for i in range(k)], side[key]

    def _forward_graphed(self, x, timesteps, context=None, class_labels=None):
        """forward() through a CUDA graph cached per (input shapes, number of timesteps, weights): the ~150 launches of
        a noise prediction become one replay; inputs are copied into the graph's static buffers (timesteps, context
        and class labels stay on the device, so any value replays).  The second call with a given signature captures,
        the first one runs eagerly (it also lets the conv kernels measure their choices)."""
        key = (tuple(x.shape), timesteps.numel(), None if context is None else tuple(context.shape),
               None if class_labels is None else class_labels.numel(), str(x.device), self.precision,
               param_signature(self))
        hit = self._fwd_graphs.get(key)
        if hit is None:
            if len(self._fwd_graphs) >= 4:
                self._fwd_graphs.clear()
            self._fwd_graphs[key] = "warm"
            return self._forward_eager(x, timesteps, context, class_labels)
        if hit == "warm":
            static_x = x.detach().to(torch.float32).clone()
            static_t = timesteps.detach().to(device=x.device, dtype=torch.float32).clone()
            static_c = None if context is None else context.detach().to(device=x.device, dtype=torch.float32).clone()
            static_l = None if class_labels is None else class_labels.detach().to(device=x.device,
                                                                                 dtype=torch.int64).clone()
            # one eager pass in the captured configuration (sub-batch streams): the conv kernels measure their choices
            # for the sub-batch shapes now - a capturing stream can only use the cost model
            self._warm_split = True
            try:
                self._forward_eager(static_x, static_t, static_c, static_l)
            finally:
                self._warm_split = False
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                static_out = self._forward_eager(static_x, static_t, static_c, static_l)
            hit = (g, static_x, static_t, static_c, static_l, static_out)
            self._fwd_graphs[key] = hit
        g, static_x, static_t, static_c, static_l, static_out = hit
        static_x.copy_(x)
        static_t.copy_(timesteps)
        if static_c is not None:
            static_c.copy_(context)
        if static_l is not None:
            static_l.copy_(class_labels)
        g.replay()
        return static_out.clone()


# ------------------------------------------------------------------------------------------
# AutoencoderKL
# ------------------------------------------------------------------------------------------
class AEResBlock(nn.Module):
    def __init__(self, cin, cout, groups, eps):
        super().__init__()
        self.cout, self.groups = cout, groups
        self.norm1 = GroupNorm(groups, cin, eps)
        self.conv1 = ConvOnly(cin, cout)
        self.norm2 = GroupNorm(groups, cout, eps)
        self.conv2 = ConvOnly(cout, cout)
        self.nin_shortcut = nn.Identity() if cin == cout else ConvOnly(cin, cout, 1, 1, 0)

    def run(self, x):
        h = self.norm1.run(x, silu=True)
        stats = None
        cpg = self.cout // self.groups
        if FUSE_GN_STATS and cpg >= 2 and (cpg & (cpg - 1)) == 0:
            stats = ops.zeros_f64(x.shape[0] * self.groups * 2, x.device)
        h = self.conv1.run(h, gn_stats=stats, gn_groups=self.groups, want_stats=stats is None)
        h = self.norm2.run(h, silu=True, stats=stats)
        if isinstance(self.nin_shortcut, nn.Identity):
            return self.conv2.run(h, residual=x, want_stats=True)
        if FUSE_SKIP and h.dtype == torch.bfloat16:
            return self.conv2.conv.run_with_skip(h, self.nin_shortcut.conv, x, want_stats=True)
        return self.conv2.run(h, residual=self.nin_shortcut.run(x), want_stats=True)


class AEDownsample(nn.Module):
    """F.pad(x, (0,1)*3) then conv3 stride 2 pad 0 == conv with pad_lo 0, pad_hi 1 (zero fill by TMA)."""

    def __init__(self, ch):
        super().__init__()
        self.conv = ConvOnly(ch, ch, 3, 2, 0, 1)

    def run(self, x, want_stats=False):
        return self.conv.run(x, want_stats=want_stats)


class _AEStack(nn.Module):
    def run(self, x):
        for b in self.blocks:
            if isinstance(b, GroupNorm):
                x = b.run(x, silu=False)
            elif isinstance(b, (ConvOnly, AEDownsample)):
                x = b.run(x, want_stats=True)
            else:
                x = b.run(x)
        return x


class Encoder(_AEStack):
    def __init__(self, in_channels, channels, out_channels, num_res_blocks, groups, eps, attention_levels,
                 with_nonlocal_attn):
        super().__init__()
        blocks: list[nn.Module] = [ConvOnly(in_channels, channels[0])]
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
        blocks.append(GroupNorm(groups, channels[-1], eps))
        blocks.append(ConvOnly(channels[-1], out_channels))
        self.blocks = nn.ModuleList(blocks)


class Decoder(_AEStack):
    def __init__(self, channels, in_channels, out_channels, num_res_blocks, groups, eps, attention_levels,
                 with_nonlocal_attn):
        super().__init__()
        rc = list(reversed(channels))
        ra, rn = list(reversed(attention_levels)), list(reversed(num_res_blocks))
        blocks: list[nn.Module] = [ConvOnly(in_channels, rc[0])]
        if with_nonlocal_attn:
            blocks += [AEResBlock(rc[0], rc[0], groups, eps), SpatialAttentionBlock(rc[0], None, groups, eps),
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
                blocks.append(NearestUp(cin))
        blocks.append(GroupNorm(groups, cin, eps))
        blocks.append(ConvOnly(cin, out_channels, terminal=True))
        self.blocks = nn.ModuleList(blocks)


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
        precision: str | None = None,
    ):
        super().__init__()
        if spatial_dims != 3:
            raise ValueError("b200ldm implements the 3-D hot path only (spatial_dims=3)")
        if use_convtranspose:
            raise NotImplementedError("use_convtranspose is not used by any reference config")
        if any(c % norm_num_groups for c in channels):
            raise ValueError("AutoencoderKL expects all channels being multiple of norm_num_groups")
        if len(channels) != len(attention_levels):
            raise ValueError("AutoencoderKL expects channels being same size of attention_levels")
        if any(c % 8 for c in channels):
            raise NotImplementedError("b200ldm AutoencoderKL needs channel counts that are multiples of 8 (levels narrower "
                                      "than 64 channels run zero-padded to one 128-byte TMA row)")
        nres = _per_level(num_res_blocks, len(channels))
        if len(nres) != len(channels):
            raise ValueError("`num_res_blocks` should be a single integer or a tuple of integers with the same length "
                             "as `channels`.")
        self.in_channels, self.out_channels, self.latent_channels = in_channels, out_channels, latent_channels
        self.precision = precision or default_precision()
        self.encoder = Encoder(in_channels, channels, latent_channels, nres, norm_num_groups, norm_eps,
                               attention_levels, with_encoder_nonlocal_attn)
        self._decode_graphs = {}
        self.decoder = Decoder(channels, latent_channels, out_channels, nres, norm_num_groups, norm_eps,
                               attention_levels, with_decoder_nonlocal_attn)
        self.quant_conv_mu = ConvOnly(latent_channels, latent_channels, 1, 1, 0, terminal=True)
        self.quant_conv_log_sigma = ConvOnly(latent_channels, latent_channels, 1, 1, 0, terminal=True)
        self.post_quant_conv = ConvOnly(latent_channels, latent_channels, 1, 1, 0)
        for m in self.modules():   # every conv here feeds a GroupNorm with norm_num_groups groups (or nothing)
            if isinstance(m, Conv):
                m.stats_groups = norm_num_groups

    def load_old_state_dict(self, old_state_dict: dict, verbose: bool = False) -> None:
        load_old_state_dict(self, old_state_dict, verbose)

    # ---- NDHWC cores
    def decode_ndhwc(self, z):
        """z: [N,d,h,w,pad64(latent)] -> [N,4d,4h,4w,round16(out_channels)]"""
        return self.decoder.run(self.post_quant_conv.run(z))

    def encode_ndhwc(self, x):
        h = self.encoder.run(x)
        return self.quant_conv_mu.run(h), self.quant_conv_log_sigma.run(h)

    # ---- MONAI surface (NCDHW fp32)
    @torch.no_grad()
    def encode(self, x):
        dt = _PREC[self.precision]
        # a tap-packed conv_in reads its 1 or 2 input channels with any row stride: 8 channels (16 bytes) instead of 64
        narrow = self.encoder.blocks[0].conv.tap_packed(torch.empty(0, dtype=dt))
        mu, logvar = self.encode_ndhwc(ops.to_ndhwc(x, 8 if narrow else _pad64(self.in_channels), dt))
        z, sigma = ops.ae_sampling(mu, logvar, None)  # z == mu (fp32), sigma = exp(clamp(logvar)/2)
        return ops.to_ncdhw(z, self.latent_channels), ops.to_ncdhw(sigma, self.latent_channels)

    @torch.no_grad()
    def sampling(self, z_mu, z_sigma, noise=None):
        """z = mu + sigma * eps.  `noise` (optional, parity tests) replaces the internally drawn eps."""
        if noise is None:
            n = z_mu.numel()
            noise = ops.philox_normal(n, int(torch.randint(0, 2 ** 31 - 1, (1,)).item()), 0, z_mu.device)
            noise = noise.reshape(z_mu.shape)
        return ops.mul_add_f32(z_mu, z_sigma, noise)

    @torch.no_grad()
    def decode(self, z):
        dt = _PREC[self.precision]
        y = self.decode_ndhwc(ops.to_ndhwc(z, _pad64(self.latent_channels), dt))
        return ops.to_ncdhw(y, self.out_channels)

    def reconstruct(self, x):
        return self.decode(self.encode(x)[0])

    def forward(self, x):
        z_mu, z_sigma = self.encode(x)
        return self.decode(self.sampling(z_mu, z_sigma)), z_mu, z_sigma

    def encode_stage_2_inputs(self, x, noise=None):
        z_mu, z_sigma = self.encode(x)
        return self.sampling(z_mu, z_sigma, noise)

    def decode_stage_2_outputs(self, z):
        return self.decode(z)

    @torch.no_grad()
    def decode_graphed(self, z):
        """decode(z) through a CUDA graph cached per (shape, weights): the sampling hot path's decode.  Replays reuse the
        graph's private activation pool, so there is no allocator traffic after the first call."""
        key = (tuple(z.shape), str(z.device), self.precision, param_signature(self))
        hit = self._decode_graphs.get(key)
        if hit is None:
            self._decode_graphs.clear()
            static_in = z.detach().to(torch.float32).clone()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self.decode(static_in)      # warm-up: weight repack, cudaFuncSetAttribute
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                static_out = self.decode(static_in)
            hit = (g, static_in, static_out)
            self._decode_graphs[key] = hit
        g, static_in, static_out = hit
        static_in.copy_(z)
        g.replay()
        return static_out.clone()

