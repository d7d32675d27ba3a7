"""Thin tensor-level wrappers over the C ABI (include/b200ldm.h).

PyTorch is used for device memory (torch.empty), the current stream and nothing else:
every arithmetic op below is a libb200ldm kernel.  Activations are NDHWC tensors
`[N, D, H, W, C]`, contiguous, bf16 (tensor-core path) or fp32 (validation path).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import BF16, F32, ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05  # noqa: F401


class KernelProfiler:
    """CUDA-event timer around individual kernel launches on the launching stream (bench.py roofline leg).
    Enabled by setting `ops.PROFILER = KernelProfiler()`; eager mode only (events are not graph-capturable)."""

    def __init__(self):
        self.records = []  # (kind, tag, work, start_event, stop_event)

    def start(self):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream())
        return ev

    def stop(self, kind, tag, work, ev0):
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record(torch.cuda.current_stream())
        self.records.append((kind, tag, work, ev0, ev1))

    def summary(self):
        """kind -> dict(launches, ms, work); per (kind, tag) too.  Call after torch.cuda.synchronize()."""
        out, by_tag = {}, {}
        for kind, tag, work, e0, e1 in self.records:
            ms = e0.elapsed_time(e1)
            for d, k in ((out, kind), (by_tag, (kind, tag))):
                s = d.setdefault(k, {"launches": 0, "ms": 0.0, "work": 0.0})
                s["launches"] += 1
                s["ms"] += ms
                s["work"] += work
        return out, by_tag


PROFILER: KernelProfiler | None = None
# GroupNorm statistics fused into the producing conv's epilogue (per-8-channel chunk sums)
FUSE_GN_STATS = __import__("os").environ.get("B200LDM_FUSE_GN_STATS", "1") != "0"
# conv kernel choice measured on first use (b200ldm_conv3d_tune); B200LDM_CONV_AUTOTUNE=0 also disables it inside the library
AUTOTUNE = __import__("os").environ.get("B200LDM_CONV_AUTOTUNE", "1") != "0"
# split-K convs feeding a small-level GroupNorm leave their finalize pass to it (b200ldm_groupnorm_small_slabs)
DEFER_FINALIZE = __import__("os").environ.get("B200LDM_DEFER_FINALIZE", "1") != "0"
GN_SMALL_MAX_S = int(__import__("os").environ.get("B200LDM_GN_SMALL_MAX_S", "256"))


def dtype_code(t: torch.dtype) -> int:
    if t == torch.bfloat16:
        return BF16
    if t == torch.float32:
        return F32
    raise TypeError(f"b200ldm supports bf16 / fp32 activations, got {t}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class PendingFinalize:
    """A split-K conv launched with defer_finalize: its output tensor `y` is allocated but NOT written yet - the fp32
    partial sums sit in `ws`.  The tensor carries this object as `y.pending` until someone sums the slabs: the GroupNorm
    that consumes it (groupnorm -> b200ldm_groupnorm_small_slabs, one launch for finalize + statistics + normalise), or
    `_ptr`, which every other op goes through and which runs the standalone finalize first."""

    def __init__(self, ws, splits, bias, temb, temb_stride, residual, M, cout, vox_per_sample):
        self.ws, self.splits, self.bias, self.temb, self.temb_stride = ws, splits, bias, temb, temb_stride
        self.residual, self.M, self.cout, self.vox_per_sample = residual, M, cout, vox_per_sample


def finalize(t: torch.Tensor) -> None:
    pend = getattr(t, "pending", None)
    if pend is None:
        return
    t.pending = None
    _lib.check(_lib.load().b200ldm_conv3d_finalize(pend.ws.data_ptr(), pend.splits, pend.M, pend.cout,
                                                   pend.vox_per_sample, _ptr(pend.bias), _ptr(pend.temb),
                                                   pend.temb_stride, _ptr(pend.residual), t.data_ptr(), _stream()),
               "conv3d_finalize")


def _ptr(t) -> int | None:
    if t is None:
        return None
    if getattr(t, "pending", None) is not None:
        finalize(t)
    return t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("b200ldm kernels need CUDA tensors (there is no CPU fallback)")


def round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def conv_out_extent(n: int, k: int, stride: int, pad_lo: int, pad_hi: int) -> int:
    return (n + pad_lo + pad_hi - k) // stride + 1


def conv3d(x, w_packed, bias, cout: int, ksize: int = 3, stride: int = 1, pad_lo: int = 1, pad_hi: int | None = None,
           x2=None, temb=None, temb_stride: int = 0, residual=None, gn_stats=None, gn_groups: int = 0,
           out_f32: bool = False, engine: int = ENGINE_AUTO, upsample2x: bool = False, want_stats: bool = False,
           skip=None, out=None, cstats_out=None, stats_width: int = 8, defer: bool = False):
    """y = conv3d(cat(x, x2)) + bias (+ temb) (+ residual); x: [N,D,H,W,C1] NDHWC.
    upsample2x: the conv runs on the nearest-x2-upsampled x without materialising it (bf16: phase-decomposed
    weights from repack_upconv_weight; fp32: ordinary weights).
    skip: (sx, sx2 | None, sw_packed): a 1x1x1 conv of cat(sx, sx2) accumulated into the same output (ResNet
    skip_connection fused as extra K iterations; `bias` must already include its bias).  bf16 only.
    want_stats: when the tensor-core epilogue can, it also accumulates per-8-channel (sum, sumsq) of y; the result is
    attached as `y.cstats` and lets the following GroupNorm skip its statistics pass.
    stats_width: channels per statistics entry (8; 2 or 4 = the group width of a following GroupNorm with fewer than
    8 channels per group - the persistent kernels emit those for Cout tiles <= 128); attached as `y.cstats_width`.
    out / cstats_out: write into this (contiguous, e.g. a batch slice of a larger tensor) instead of allocating; the
    statistics slice must arrive zeroed.
    defer: the caller hands `y` straight to a GroupNorm (ops.groupnorm) - when the problem runs split-K the finalize
    pass is left to that GroupNorm (see PendingFinalize); the returned tensor must not be viewed / sliced before."""
    _need_cuda(x, x2, w_packed, bias, temb, residual)
    lib = _lib.load()
    N, D, H, W, C1 = x.shape
    C2 = 0 if x2 is None else x2.shape[-1]
    if pad_hi is None:
        pad_hi = pad_lo
    up = 2 if upsample2x else 1
    Do, Ho, Wo = (conv_out_extent(up * v, ksize, stride, pad_lo, pad_hi) for v in (D, H, W))
    odt = torch.float32 if out_f32 else x.dtype
    if out is None:
        y = torch.empty((N, Do, Ho, Wo, cout), dtype=odt, device=x.device)
    else:
        assert tuple(out.shape) == (N, Do, Ho, Wo, cout) and out.dtype == odt and out.is_contiguous() and out.is_cuda
        y = out
    p = _lib.ConvParams(
        x=_ptr(x), x2=_ptr(x2), w=_ptr(w_packed), bias=_ptr(bias), temb=_ptr(temb), residual=_ptr(residual),
        y=_ptr(y), gn_stats=_ptr(gn_stats), N=N, D=D, H=H, W=W, C1=C1, C2=C2, Cout=cout, Do=Do, Ho=Ho, Wo=Wo,
        ksize=ksize, stride=stride, pad_lo=pad_lo, temb_stride=temb_stride, gn_groups=gn_groups,
        dtype=dtype_code(x.dtype), engine=engine, out_f32=int(out_f32), upsample2x=int(upsample2x),
        chunk_width=stats_width if want_stats else 0)
    if skip is not None:
        sx, sx2, sw = skip
        _need_cuda(sx, sx2, sw)
        assert sx.is_contiguous() and tuple(sx.shape[:-1]) == (N, Do, Ho, Wo) and (sx2 is None or sx2.is_contiguous())
        p.skip_x, p.skip_x2, p.skip_w = _ptr(sx), _ptr(sx2), _ptr(sw)
        p.skip_C1, p.skip_C2 = sx.shape[-1], 0 if sx2 is None else sx2.shape[-1]
    assert x.is_contiguous() and (x2 is None or x2.is_contiguous())
    assert residual is None or (residual.is_contiguous() and residual.shape == y.shape and residual.dtype == odt)
    if AUTOTUNE and lib.b200ldm_conv3d_tune_pending(C.byref(p)) and not torch.cuda.is_current_stream_capturing():
        # first eager call of this problem signature: measure which kernel / split-K factor runs it fastest (the one
        # synchronising entry point of the library, include/b200ldm.h); b200ldm_conv3d below - and every later call,
        # captured or not - then runs the remembered choice and never synchronises
        tws = lib.b200ldm_conv3d_workspace_bytes(C.byref(p))      # the largest candidate while tuning is pending
        if tws > 0:
            ws = torch.empty((tws,), dtype=torch.uint8, device=x.device)
            p.workspace, p.workspace_bytes = ws.data_ptr(), tws
        if want_stats and FUSE_GN_STATS:
            scratch = zeros_f64(N * (cout // stats_width) * 2, x.device)
            p.chunk_stats = scratch.data_ptr()
        _lib.check(lib.b200ldm_conv3d_tune(C.byref(p), _stream()), "conv3d_tune")
        p.workspace, p.workspace_bytes, p.chunk_stats = None, 0, None
    ws_bytes = lib.b200ldm_conv3d_workspace_bytes(C.byref(p))
    ws = None
    if ws_bytes > 0:  # split-K partial sums (small-M convs): caller-owned scratch
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x.device)
        p.workspace, p.workspace_bytes = ws.data_ptr(), ws_bytes
    splits = 1
    if defer and ws is not None and out is None and not out_f32 and DEFER_FINALIZE and Do * Ho * Wo <= GN_SMALL_MAX_S:
        # (larger tensors: the finalize pass of the conv call itself also accumulates the GroupNorm chunk sums)
        splits = lib.b200ldm_conv3d_splits(C.byref(p))
        p.defer_finalize = 1 if splits > 1 else 0
    cstats = None
    if want_stats and FUSE_GN_STATS and lib.b200ldm_conv3d_emits_chunk_stats(C.byref(p)):
        if cstats_out is not None:
            assert cstats_out.dtype == torch.float64 and cstats_out.numel() == N * (cout // stats_width) * 2
            cstats = cstats_out
        else:
            cstats = zeros_f64(N * (cout // stats_width) * 2, x.device)
        p.chunk_stats = cstats.data_ptr()
    ev = PROFILER.start() if PROFILER is not None else None
    _lib.check(lib.b200ldm_conv3d(C.byref(p), _stream()), "conv3d")
    if p.defer_finalize:
        y.pending = PendingFinalize(ws, splits, bias, temb, temb_stride, residual, N * Do * Ho * Wo, cout, Do * Ho * Wo)
    if cstats is not None:
        y.cstats, y.cstats_width = cstats, stats_width
    if ev is not None:
        # algorithmic FLOPs of the (padded) problem the kernel ran: 2*M*Cout*Cin*k^3
        # (fused upsample on tensor cores: 8 of the 27 taps are executed per output voxel - EXECUTED flops)
        taps = 8 if (upsample2x and x.dtype == torch.bfloat16) else ksize ** 3
        kk = (C1 + C2) * taps + (0 if skip is None else p.skip_C1 + p.skip_C2)
        PROFILER.stop("conv3d", f"{C1 + C2}->{cout} k{ksize}s{stride}{'up' if upsample2x else ''}"
                                f"{'+skip' if skip is not None else ''} @{N}x{Do}x{Ho}x{Wo}",
                      2.0 * N * Do * Ho * Wo * cout * kk, ev)
    return y


def linear(x2d, w_packed, bias, cout: int, residual=None, out_f32: bool = False, engine: int = ENGINE_AUTO):
    """[M, K] @ W^T (+bias) (+residual) as a 1x1x1 conv over a [1,1,1,M,K] view."""
    M, K = x2d.shape
    r = None if residual is None else residual.reshape(1, 1, 1, M, cout)
    y = conv3d(x2d.reshape(1, 1, 1, M, K), w_packed, bias, cout, ksize=1, stride=1, pad_lo=0, residual=r,
               out_f32=out_f32, engine=engine)
    return y.reshape(M, cout)


def repack_conv_weight(w: torch.Tensor, cout_pad: int, cin_pad: int, dtype: torch.dtype, cin_split: int | None = None,
                       cin_split_pad: int | None = None):
    """fp32 [Cout,Cin,k,k,k] (or [Cout,Cin] for linears) -> packed [k^3][cout_pad][cin_pad]."""
    _need_cuda(w)
    lib = _lib.load()
    w = w.detach().to(torch.float32).contiguous()
    cout, cin = w.shape[0], w.shape[1]
    ksize = w.shape[2] if w.ndim == 5 else 1
    if cin_split is None:
        cin_split, cin_split_pad = cin, cin_pad
    out = torch.empty((ksize ** 3, cout_pad, cin_pad), dtype=dtype, device=w.device)
    _lib.check(lib.b200ldm_repack_conv_weight(w.data_ptr(), out.data_ptr(), cout, cin, ksize, cout_pad, cin_pad,
                                              cin_split, cin_split_pad, dtype_code(dtype), _stream()), "repack")
    return out


def repack_upconv_weight(w: torch.Tensor, cout_pad: int, cin_pad: int, dtype: torch.dtype):
    """fp32 [Cout,Cin,3,3,3] -> [64 = 8 phases x 8 taps][cout_pad][cin_pad] for the fused nearest-x2 upsample conv."""
    _need_cuda(w)
    w = w.detach().to(torch.float32).contiguous()
    assert w.ndim == 5 and tuple(w.shape[2:]) == (3, 3, 3)
    out = torch.empty((64, cout_pad, cin_pad), dtype=dtype, device=w.device)
    _lib.check(_lib.load().b200ldm_repack_upconv_weight(w.data_ptr(), out.data_ptr(), w.shape[0], w.shape[1], cout_pad,
                                                        cin_pad, dtype_code(dtype), _stream()), "repack_upconv")
    return out


def zeros_f64(n: int, device) -> torch.Tensor:
    t = torch.empty((n,), dtype=torch.float64, device=device)
    _lib.check(_lib.load().b200ldm_fill_zero(t.data_ptr(), n * 8, _stream()), "fill_zero")
    return t


def groupnorm(x, gamma, beta, groups: int, eps: float, silu: bool, x2=None, stats=None):
    """GroupNorm(+SiLU) over cat(x, x2) channels; `stats` may come pre-accumulated from a conv epilogue."""
    _need_cuda(x, x2, gamma, beta)
    lib = _lib.load()
    N = x.shape[0]
    C1 = x.shape[-1]
    C2 = 0 if x2 is None else x2.shape[-1]
    S = x.numel() // (N * C1)
    dt = dtype_code(x.dtype)
    nbytes = float(N * S * (C1 + C2) * x.element_size())
    pend = getattr(x, "pending", None)
    if pend is not None:
        cpg = (C1 + C2) // groups
        if (stats is None and x.dtype == torch.bfloat16 and S <= GN_SMALL_MAX_S and (C1 + C2) % groups == 0
                and cpg % 8 == 0 and C1 % cpg == 0 and S * (cpg // 8) <= 2048):
            # the producing split-K conv left its partial sums: finalize + statistics + normalise in ONE launch; it also
            # writes the un-normalised tensor into x itself (later residual / skip reads see a finished tensor)
            x.pending = None
            y = torch.empty(x.shape[:-1] + (C1 + C2,), dtype=x.dtype, device=x.device)
            ev = PROFILER.start() if PROFILER is not None else None
            _lib.check(lib.b200ldm_groupnorm_small_slabs(
                pend.ws.data_ptr(), pend.splits, _ptr(pend.bias), _ptr(pend.temb), pend.temb_stride, _ptr(pend.residual),
                x.data_ptr(), _ptr(x2), y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), N, S, C1, C2, groups,
                float(eps), int(silu), _stream()), "groupnorm_small_slabs")
            if ev is not None:
                PROFILER.stop("gn_small", f"C{C1 + C2} S{S} +finalize x{pend.splits}",
                              2.0 * nbytes + float(N * S * C1) * (4 * pend.splits + 2), ev)
            return y
        finalize(x)
    cs1, cs2 = getattr(x, "cstats", None), (None if x2 is None else getattr(x2, "cstats", None))
    cw = getattr(x, "cstats_width", 8)
    if (stats is None and cs1 is not None and (x2 is None or (cs2 is not None and getattr(x2, "cstats_width", 8) == cw))
            and ((C1 + C2) // groups) % cw == 0 and C1 % cw == 0):
        # statistics already accumulated by the producing conv(s): one pass only
        y = torch.empty(x.shape[:-1] + (C1 + C2,), dtype=x.dtype, device=x.device)
        ev = PROFILER.start() if PROFILER is not None else None
        _lib.check(lib.b200ldm_groupnorm_apply_chunked_w(_ptr(x), _ptr(x2), y.data_ptr(), cs1.data_ptr(), _ptr(cs2),
                                                         gamma.data_ptr(), beta.data_ptr(), N, S, C1, C2, groups,
                                                         float(eps), int(silu), dt, cw, _stream()),
                   "groupnorm_apply_chunked")
        if ev is not None:
            PROFILER.stop("gn_apply", f"C{C1 + C2} S{S}", 2.0 * nbytes, ev)
        return y
    if stats is None and S <= GN_SMALL_MAX_S and lib.b200ldm_groupnorm_small_fits(S, C1 + C2, groups, dt):
        # small levels: one launch instead of memset + stats + apply
        y = torch.empty(x.shape[:-1] + (C1 + C2,), dtype=x.dtype, device=x.device)
        ev = PROFILER.start() if PROFILER is not None else None
        _lib.check(lib.b200ldm_groupnorm_small(_ptr(x), _ptr(x2), y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), N, S,
                                               C1, C2, groups, float(eps), int(silu), dt, _stream()), "groupnorm_small")
        if ev is not None:
            PROFILER.stop("gn_small", f"C{C1 + C2} S{S}", 2.0 * nbytes, ev)
        return y
    if stats is None:
        stats = zeros_f64(N * groups * 2, x.device)
        ev = PROFILER.start() if PROFILER is not None else None
        _lib.check(lib.b200ldm_groupnorm_stats(_ptr(x), _ptr(x2), stats.data_ptr(), N, S, C1, C2, groups, dt,
                                               _stream()), "groupnorm_stats")
        if ev is not None:
            PROFILER.stop("gn_stats", f"C{C1 + C2} S{S}", nbytes, ev)  # one read
    y = torch.empty(x.shape[:-1] + (C1 + C2,), dtype=x.dtype, device=x.device)
    ev = PROFILER.start() if PROFILER is not None else None
    _lib.check(lib.b200ldm_groupnorm_apply(_ptr(x), _ptr(x2), y.data_ptr(), stats.data_ptr(), gamma.data_ptr(),
                                           beta.data_ptr(), N, S, C1, C2, groups, float(eps), int(silu), dt,
                                           _stream()), "groupnorm_apply")
    if ev is not None:
        PROFILER.stop("gn_apply", f"C{C1 + C2} S{S}", 2.0 * nbytes, ev)  # one read + one write
    return y


def attention(qkv, heads: int, head_dim: int, scale: float, engine: int = ENGINE_AUTO):
    """qkv: [B, L, 3*H*d] fused projections -> [B, L, H*d]."""
    _need_cuda(qkv)
    lib = _lib.load()
    B, L, C3 = qkv.shape
    Cc = heads * head_dim
    assert C3 == 3 * Cc and qkv.is_contiguous()
    o = torch.empty((B, L, Cc), dtype=qkv.dtype, device=qkv.device)
    es = qkv.element_size()
    base = qkv.data_ptr()
    # scratch for splitting the key range of the last, partial wave of query tiles (0 when that gains nothing)
    ws_bytes = lib.b200ldm_attention_workspace_bytes(B, heads, L, L, head_dim, dtype_code(qkv.dtype), engine)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=qkv.device) if ws_bytes > 0 else None
    ev = PROFILER.start() if PROFILER is not None else None
    _lib.check(lib.b200ldm_attention_ws(base, base + Cc * es, base + 2 * Cc * es, o.data_ptr(), B, heads, L, head_dim,
                                        C3, float(scale), dtype_code(qkv.dtype), engine, _ptr(ws), ws_bytes, _stream()),
               "attention")
    if ev is not None:
        PROFILER.stop("attention", f"B{B} H{heads} L{L} d{head_dim}", 4.0 * B * heads * L * L * head_dim, ev)
    return o


def cross_attention(q, kv, heads: int, head_dim: int, scale: float, engine: int = ENGINE_AUTO):
    """q: [B, Lq, H*d]; kv: [B, Lk, 2*H*d] fused key/value projections of the context -> [B, Lq, H*d]."""
    _need_cuda(q, kv)
    lib = _lib.load()
    B, Lq, Cc = q.shape
    Lk = kv.shape[1]
    assert Cc == heads * head_dim and kv.shape == (B, Lk, 2 * Cc) and q.is_contiguous() and kv.is_contiguous()
    assert q.dtype == kv.dtype
    o = torch.empty((B, Lq, Cc), dtype=q.dtype, device=q.device)
    base = kv.data_ptr()
    ev = PROFILER.start() if PROFILER is not None else None
    _lib.check(lib.b200ldm_cross_attention(q.data_ptr(), base, base + Cc * kv.element_size(), o.data_ptr(), B, heads,
                                           Lq, Lk, head_dim, Cc, 2 * Cc, float(scale), dtype_code(q.dtype), engine,
                                           _stream()), "cross_attention")
    if ev is not None:
        PROFILER.stop("attention", f"B{B} H{heads} Lq{Lq} Lk{Lk} d{head_dim}", 4.0 * B * heads * Lq * Lk * head_dim, ev)
    return o


def layernorm(x2d, gamma, beta, eps: float = 1e-5):
    """nn.LayerNorm over the last axis of [M, C] tokens (gamma / beta fp32)."""
    _need_cuda(x2d, gamma, beta)
    M, Cc = x2d.shape
    assert x2d.is_contiguous() and gamma.dtype == torch.float32 and beta.dtype == torch.float32
    y = torch.empty_like(x2d)
    ev = PROFILER.start() if PROFILER is not None else None
    _lib.check(_lib.load().b200ldm_layernorm(x2d.data_ptr(), y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), M, Cc,
                                             float(eps), dtype_code(x2d.dtype), _stream()), "layernorm")
    if ev is not None:
        PROFILER.stop("layernorm", f"C{Cc} M{M}", 2.0 * x2d.numel() * x2d.element_size(), ev)
    return y


def geglu(h2d):
    """[M, 2J] -> [M, J]: value half * gelu(gate half) (MONAI MLPBlock act="GEGLU")."""
    _need_cuda(h2d)
    M, J2 = h2d.shape
    assert h2d.is_contiguous() and J2 % 16 == 0
    y = torch.empty((M, J2 // 2), dtype=h2d.dtype, device=h2d.device)
    ev = PROFILER.start() if PROFILER is not None else None
    _lib.check(_lib.load().b200ldm_geglu(h2d.data_ptr(), y.data_ptr(), M, J2 // 2, dtype_code(h2d.dtype), _stream()),
               "geglu")
    if ev is not None:
        PROFILER.stop("geglu", f"J{J2 // 2} M{M}", 1.5 * h2d.numel() * h2d.element_size(), ev)
    return y


def embedding_add(a, table, idx):
    """a [n, dim] fp32 + table[idx] (class embedding onto the time embedding); idx int64 on the device."""
    _need_cuda(a, table, idx)
    n, dim = a.shape
    assert a.dtype == torch.float32 and table.dtype == torch.float32 and idx.dtype == torch.int64
    assert a.is_contiguous() and table.is_contiguous() and idx.numel() == n and table.shape[1] == dim
    out = torch.empty_like(a)
    _lib.check(_lib.load().b200ldm_embedding_add(a.data_ptr(), table.data_ptr(), idx.data_ptr(), out.data_ptr(), n, dim,
                                                 table.shape[0], _stream()), "embedding_add")
    return out


def tap_gather(z, bias, cp: int, cy: int, sched=None, want_y: bool = True, out=None):
    """y[v][co] = bias[co] + sum_tap z[v + shift(tap)][tap*cp + co]: second half of a narrow-output 3x3x3 conv.
    sched: a SchedParams (make_sched_params) -> the gathered tensor is the noise prediction and the scheduler update is
    applied in the same kernel; with want_y=False the prediction itself is never written (returns None)."""
    _need_cuda(z, bias)
    N, D, H, W, ldz = z.shape
    assert z.is_contiguous()
    y = None
    if out is not None:
        assert tuple(out.shape) == (N, D, H, W, cy) and out.dtype == z.dtype and out.is_contiguous() and out.is_cuda
        y = out
    elif want_y or sched is None:
        y = torch.empty((N, D, H, W, cy), dtype=z.dtype, device=z.device)
    ev = PROFILER.start() if PROFILER is not None else None
    _lib.check(_lib.load().b200ldm_tap_gather(_ptr(z), _ptr(bias), _ptr(y), N, D, H, W, ldz, cp, cy,
                                              dtype_code(z.dtype), None if sched is None else C.byref(sched), _stream()),
               "tap_gather")
    if ev is not None:
        PROFILER.stop("tap_gather", f"cp{cp} S{D * H * W}{'+sched' if sched is not None else ''}",
                      float(N * D * H * W) * (27 * cp + cy) * z.element_size(), ev)
    return y


def upsample_nearest2x(x):
    _need_cuda(x)
    N, D, H, W, Cc = x.shape
    y = torch.empty((N, 2 * D, 2 * H, 2 * W, Cc), dtype=x.dtype, device=x.device)
    _lib.check(_lib.load().b200ldm_upsample_nearest2x(_ptr(x), y.data_ptr(), N, D, H, W, Cc,
                                                      dtype_code(x.dtype), _stream()), "upsample")
    return y


def pack_taps(x: torch.Tensor, cin: int) -> torch.Tensor:
    """[N,D,H,W,Cp] (first `cin` <= 2 channels real) -> [N,D,H,W,64]: channel t * cin + c = x at tap t of the 3x3x3
    neighbourhood (zero padding), so that the conv becomes ONE 64-wide k-iteration (b200ldm_pack_taps)."""
    _need_cuda(x)
    N, D, H, W, Cp = x.shape
    assert x.is_contiguous() and cin in (1, 2) and Cp >= cin
    y = torch.empty((N, D, H, W, 64), dtype=x.dtype, device=x.device)
    ev = PROFILER.start() if PROFILER is not None else None
    _lib.check(_lib.load().b200ldm_pack_taps(_ptr(x), y.data_ptr(), N, D, H, W, Cp, cin, dtype_code(x.dtype), _stream()),
               "pack_taps")
    if ev is not None:
        PROFILER.stop("pack_taps", f"cin{cin} S{D * H * W}", float(N * D * H * W) * (27 * cin + 64) * x.element_size(), ev)
    return y


def to_ndhwc(src: torch.Tensor, c_dst: int, dtype: torch.dtype, out: torch.Tensor | None = None, c_offset: int = 0):
    """NCDHW fp32 -> NDHWC `dtype`, channels [c_offset, c_offset+C) of a c_dst-wide (zero padded) buffer."""
    _need_cuda(src)
    src = src.to(torch.float32).contiguous()
    N, Cc = src.shape[:2]
    sp = tuple(src.shape[2:])
    S = src.numel() // (N * Cc)
    if out is None:
        out = torch.empty((N,) + sp + (c_dst,), dtype=dtype, device=src.device)
        if c_dst != Cc:
            _lib.check(_lib.load().b200ldm_fill_zero(out.data_ptr(), out.numel() * out.element_size(), _stream()),
                       "fill_zero")
    _lib.check(_lib.load().b200ldm_ncdhw_to_ndhwc(src.data_ptr(), out.data_ptr(), N, Cc, S, c_dst, c_offset,
                                                  dtype_code(out.dtype), _stream()), "ncdhw_to_ndhwc")
    return out


def to_ncdhw(src: torch.Tensor, channels: int) -> torch.Tensor:
    """NDHWC (bf16|fp32, possibly channel padded) -> NCDHW fp32 with the first `channels` channels."""
    _need_cuda(src)
    N = src.shape[0]
    sp = tuple(src.shape[1:-1])
    c_src = src.shape[-1]
    S = src.numel() // (N * c_src)
    out = torch.empty((N, channels) + sp, dtype=torch.float32, device=src.device)
    _lib.check(_lib.load().b200ldm_ndhwc_to_ncdhw(_ptr(src), out.data_ptr(), N, channels, S, c_src,
                                                  dtype_code(src.dtype), _stream()), "ndhwc_to_ncdhw")
    return out


def silu_f32(x):
    _need_cuda(x)
    y = torch.empty_like(x)
    _lib.check(_lib.load().b200ldm_silu_f32(x.data_ptr(), y.data_ptr(), x.numel(), _stream()), "silu")
    return y


def timestep_embedding(timesteps: torch.Tensor, dim: int, max_period: float = 10000.0):
    _need_cuda(timesteps)
    ts = timesteps.to(torch.float32).contiguous()
    out = torch.empty((ts.numel(), dim), dtype=torch.float32, device=ts.device)
    _lib.check(_lib.load().b200ldm_timestep_embedding(ts.data_ptr(), out.data_ptr(), ts.numel(), dim,
                                                      float(max_period), _stream()), "timestep_embedding")
    return out


def select_row(table, counter, dst):
    _lib.check(_lib.load().b200ldm_select_row(table.data_ptr(), counter.data_ptr(), dst.data_ptr(), dst.numel(),
                                              _stream()), "select_row")


def counter_add(counter, delta: int = 1):
    _lib.check(_lib.load().b200ldm_counter_add(counter.data_ptr(), delta, _stream()), "counter_add")


def make_sched_params(x_t, coef, counter, c: int, dtype: torch.dtype, eps=None, noise=None, noise_step_stride: int = 0,
                      x0_out=None, x_model=None, seed=0, var_offset: int = 0, var_coef=None, philox_offset: int = 0):
    """b200ldm_sched_params for `scheduler_step` or for the fused update inside `tap_gather`.  The returned struct holds
    raw pointers: keep the tensors alive (the sampling loop owns them for its whole life).
    var_coef ([n_steps, 2] fp32) + var_offset: learned variance, the predicted variance sits `var_offset` channels
    into `eps` (the second half of the model output)."""
    _need_cuda(eps, x_t, coef, noise, x_model, var_coef)
    seed_dev = None
    if torch.is_tensor(seed):
        _need_cuda(seed)
        assert seed.dtype == torch.int64 and seed.numel() == 1
        seed_dev, seed = seed, 0
    p = _lib.SchedParams(
        eps=_ptr(eps), x_t=_ptr(x_t), noise=_ptr(noise), x0_out=_ptr(x0_out), x_model=_ptr(x_model), coef=_ptr(coef),
        counter=_ptr(counter), n_vox=x_t.numel() // c, noise_step_stride=noise_step_stride, C=c,
        C_eps=c if eps is None else eps.shape[-1], C_model=0 if x_model is None else x_model.shape[-1],
        dtype=dtype_code(dtype), seed=seed, seed_dev=_ptr(seed_dev), philox_offset=philox_offset)
    if var_coef is not None:
        assert eps is not None and 0 < var_offset and var_offset + c <= eps.shape[-1] and var_coef.dtype == torch.float32
        p.var_pred, p.var_coef = eps.data_ptr() + var_offset * eps.element_size(), var_coef.data_ptr()
    p._keep = (eps, x_t, noise, x0_out, x_model, coef, counter, seed_dev, var_coef)
    return p


def scheduler_step(eps, x_t, coef, counter, c: int, noise=None, noise_step_stride: int = 0, x0_out=None, x_model=None,
                   seed=0, var_offset: int = 0, var_coef=None, philox_offset: int = 0):
    """In-place x_t <- step(eps, x_t); eps [.., C_eps] NDHWC, x_t fp32 [.., C] NDHWC.
    seed: int, or a device int64[1] tensor read by the kernel (a captured graph then serves any seed)."""
    p = make_sched_params(x_t, coef, counter, c, eps.dtype, eps=eps, noise=noise, noise_step_stride=noise_step_stride,
                          x0_out=x0_out, x_model=x_model, seed=seed, var_offset=var_offset, var_coef=var_coef,
                          philox_offset=philox_offset)
    _lib.check(_lib.load().b200ldm_scheduler_step(C.byref(p), _stream()), "scheduler_step")


def philox_normal(n: int, seed: int, step: int, device) -> torch.Tensor:
    out = torch.empty((n,), dtype=torch.float32, device=device)
    _lib.check(_lib.load().b200ldm_philox_normal(out.data_ptr(), n, seed, step, _stream()), "philox_normal")
    return out


def ae_sampling(mu, logvar, noise):
    _need_cuda(mu, logvar, noise)
    z = torch.empty(mu.shape, dtype=torch.float32, device=mu.device)
    sig = torch.empty(mu.shape, dtype=torch.float32, device=mu.device)
    _lib.check(_lib.load().b200ldm_ae_sampling(mu.data_ptr(), logvar.data_ptr(), _ptr(noise), z.data_ptr(),
                                               sig.data_ptr(), mu.numel(), dtype_code(mu.dtype), _stream()),
               "ae_sampling")
    return z, sig


def add_noise(x0, noise, sqrt_acp, sqrt_1m_acp):
    _need_cuda(x0, noise, sqrt_acp, sqrt_1m_acp)
    out = torch.empty_like(x0)
    N = x0.shape[0]
    _lib.check(_lib.load().b200ldm_add_noise(x0.data_ptr(), noise.data_ptr(), sqrt_acp.data_ptr(),
                                             sqrt_1m_acp.data_ptr(), out.data_ptr(), N, x0.numel() // N, _stream()),
               "add_noise")
    return out


def mul_add_f32(a, b, c):
    """a + b * c elementwise, fp32 (AutoencoderKL.sampling: mu + sigma * eps)."""
    _need_cuda(a, b, c)
    a, b, c = (t.to(torch.float32).contiguous() for t in (a, b, c))
    out = torch.empty_like(a)
    _lib.check(_lib.load().b200ldm_mul_add_f32(a.data_ptr(), b.data_ptr(), c.data_ptr(), out.data_ptr(), a.numel(),
                                               _stream()), "mul_add_f32")
    return out


def launch_count() -> int:
    return int(_lib.load().b200ldm_launch_count())
