"""GPU parity tests of every kernel behind the C ABI against plain fp32 torch math (the oracle's
building blocks: F.conv3d / F.group_norm / softmax / the oracle schedulers), on the same seeded inputs.

Tolerances: max|a-b| / max|b| <= 1e-4 for the fp32 validation engine and <= 2e-2 for the
bf16 tensor-core engine.  For bf16 the reference is evaluated in fp64 on the bf16-ROUNDED operands, so
the only differences left are accumulation order and the final bf16 rounding (tight 1e-2 bound).
"""
import math
import os

import pytest
import torch
import torch.nn.functional as F

import golden_cases as gc
import oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = {torch.float32: 1e-4, torch.bfloat16: 1e-2}


def relerr(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def rnd(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


def q(t, dtype):
    """round to the activation dtype and come back to fp32 (what the kernel actually sees)"""
    return t.to(dtype).float()


def to_act(t_ncdhw, dtype, cpad=None):
    from b200ldm import ops
    c = t_ncdhw.shape[1]
    return ops.to_ndhwc(t_ncdhw.to(DEV), cpad or c, dtype)


def from_act(a, c):
    from b200ldm import ops
    return ops.to_ncdhw(a, c).cpu()


# ------------------------------------------------------------------------------------------
# conv3d
# ------------------------------------------------------------------------------------------
CONV_CASES = [
    # name,            N, (D,H,W),     C1,  C2, Cout, k, s, pad_lo, pad_hi, temb, resid
    ("k3_basic",       1, (8, 8, 8),    64,   0,  64, 3, 1, 1, 1, False, False),
    ("k3_fused_epi",   2, (6, 8, 4),   128,   0, 256, 3, 1, 1, 1, True, True),
    ("k3_two_source",  1, (5, 5, 5),    64, 128,  64, 3, 1, 1, 1, False, True),
    ("k1_linear",      1, (1, 1, 300), 256,   0, 768, 1, 1, 0, 0, False, True),
    ("k1_two_source",  2, (4, 4, 4),   128,  64, 128, 1, 1, 0, 0, False, False),
    ("k3_s2_even",     1, (8, 8, 8),    64,   0,  64, 3, 2, 1, 1, False, False),
    ("k3_s2_odd",      2, (9, 7, 6),    64,   0, 128, 3, 2, 1, 1, False, False),
    ("ae_down_even",   1, (8, 8, 8),    64,   0,  64, 3, 2, 0, 1, False, False),
    ("ae_down_odd",    1, (7, 9, 6),   128,   0, 128, 3, 2, 0, 1, False, False),
    ("cout16",         1, (6, 6, 6),    64,   0,  16, 3, 1, 1, 1, False, False),
    ("batch_in_tile",  8, (4, 4, 4),   256,   0, 256, 3, 1, 1, 1, True, False),
    ("deep_k",         2, (4, 4, 4),   512, 512, 512, 3, 1, 1, 1, False, False),
    ("odd_36_44_28",   1, (9, 11, 7),  256,   0, 512, 3, 1, 1, 1, False, False),
    ("wide_n_tiles",   1, (4, 6, 10),   64,   0, 1024, 3, 1, 1, 1, False, False),
]


def _conv_case(case, dtype, engine, with_stats=False):
    from b200ldm import ops
    name, N, (D, H, W), C1, C2, Cout, k, s, pl, phi, use_temb, use_res = case
    Cin = C1 + C2
    x = q(rnd((N, Cin, D, H, W), 1), dtype)
    w = q(rnd((Cout, Cin, k, k, k), 2, 1.0 / math.sqrt(Cin * k ** 3)), dtype)
    b = rnd((Cout,), 3, 0.1)
    xp = F.pad(x.double(), (pl, phi, pl, phi, pl, phi))
    ref = F.conv3d(xp, w.double(), b.double(), stride=s)
    temb = None
    if use_temb:
        temb = rnd((N, Cout), 4)
        ref = ref + temb.double()[:, :, None, None, None]
    res = None
    if use_res:
        res = q(rnd(tuple(ref.shape), 5), dtype)
        ref = ref + res.double()
    cout_rows = ops.round_up(Cout, 16)
    wp = ops.repack_conv_weight(w.to(DEV), cout_rows, Cin, dtype)
    xa = to_act(x[:, :C1], dtype)
    x2a = to_act(x[:, C1:], dtype) if C2 else None
    resa = to_act(res, dtype) if use_res else None
    G = 32
    stats = ops.zeros_f64(N * G * 2, DEV) if with_stats else None
    y = ops.conv3d(xa, wp, b.to(DEV), Cout, k, s, pl, phi, x2=x2a, temb=None if temb is None else temb.to(DEV),
                   temb_stride=Cout if use_temb else 0, residual=resa, gn_stats=stats, gn_groups=G if with_stats else 0,
                   engine=engine)
    torch.cuda.synchronize()
    out = from_act(y, Cout)
    assert out.shape == ref.shape, (out.shape, ref.shape)
    return out, ref.float(), (stats.cpu() if with_stats else None), y


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv3d_simt_fp32(case):
    from b200ldm import ops
    out, ref, _, _ = _conv_case(case, torch.float32, ops.ENGINE_SIMT)
    assert relerr(out, ref) <= 1e-4, relerr(out, ref)


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv3d_tcgen05_bf16(case):
    from b200ldm import ops
    out, ref, _, _ = _conv_case(case, torch.bfloat16, ops.ENGINE_TCGEN05)
    e = relerr(out, ref)
    assert e <= 1e-2, f"{case[0]}: relerr {e}"


@pytest.mark.parametrize("case", [CONV_CASES[1], CONV_CASES[10], CONV_CASES[12]], ids=["fused", "batch_tile", "odd"])
@pytest.mark.parametrize("engine_dtype", ["simt_f32", "tc_bf16"])
def test_conv3d_fused_groupnorm_partials(case, engine_dtype):
    from b200ldm import ops
    dtype, engine = (torch.float32, ops.ENGINE_SIMT) if engine_dtype == "simt_f32" else (torch.bfloat16, ops.ENGINE_TCGEN05)
    out, ref, stats, y = _conv_case(case, dtype, engine, with_stats=True)
    N, Cout, G = out.shape[0], out.shape[1], 32
    yy = out.double().reshape(N, G, -1)  # stats are taken over the values as stored
    want = torch.stack([yy.sum(-1), (yy * yy).sum(-1)], dim=-1).reshape(-1)
    assert relerr(stats, want) <= 1e-4, relerr(stats, want)


UP_CASES = [(1, (4, 4, 4), 64, 64), (2, (5, 6, 3), 128, 256), (1, (9, 11, 7), 256, 256), (8, (2, 2, 2), 64, 128)]


@pytest.mark.parametrize("case", UP_CASES, ids=[f"N{c[0]}_{'x'.join(map(str, c[1]))}_{c[2]}to{c[3]}" for c in UP_CASES])
@pytest.mark.parametrize("mode", ["tc_bf16", "simt_f32"])
def test_conv3d_fused_nearest_upsample(case, mode):
    """conv3(nearest_x2(x)): 8-phase 2x2x2 decomposition on tcgen05 (pre-summed weights), virtual upsample on SIMT."""
    from b200ldm import ops
    N, sp, Cin, Cout = case
    dtype = torch.bfloat16 if mode == "tc_bf16" else torch.float32
    x = q(rnd((N, Cin) + sp, 81), dtype)
    w = rnd((Cout, Cin, 3, 3, 3), 82, 1.0 / math.sqrt(Cin * 27))
    b = rnd((Cout,), 83, 0.1)
    if mode == "tc_bf16":
        wp = ops.repack_upconv_weight(w.to(DEV), Cout, Cin, dtype)
        # reference with the SAME pre-summed, bf16-rounded phase weights is not expressible as one conv3d; compare
        # against the exact fp64 result of the un-rounded weights with a tolerance covering the weight rounding
        tol = 1.5e-2
    else:
        wp = ops.repack_conv_weight(w.to(DEV), Cout, Cin, dtype)
        tol = 1e-4
    ref = F.conv3d(F.interpolate(x.double(), scale_factor=2.0, mode="nearest"), w.double(), b.double(), padding=1)
    y = ops.conv3d(to_act(x, dtype), wp, b.to(DEV), Cout, 3, 1, 1, 1, upsample2x=True)
    out = from_act(y, Cout)
    assert out.shape == ref.shape
    e = relerr(out, ref)
    assert e <= tol, e


STATS_CASES = [CONV_CASES[13], CONV_CASES[3],
               ("s_64_256", 2, (16, 16, 16), 64, 0, 256, 3, 1, 1, 1, False, False),
               ("s_two_src", 2, (9, 11, 7), 64, 64, 192, 3, 1, 1, 1, False, False),
               ("s_stride2", 2, (16, 12, 10), 64, 0, 128, 3, 2, 1, 1, False, False)]


@pytest.mark.parametrize("case", STATS_CASES, ids=[c[0] for c in STATS_CASES])
def test_conv3d_chunk_stats_feed_groupnorm(case):
    """conv epilogue chunk stats == sums over the stored bf16 output; GroupNorm applied from them == two-pass GroupNorm,
    including a group that straddles two concatenated sources."""
    from b200ldm import ops
    name, N, (D, H, W), C1, C2, Cout, k, s, pl, phi, use_temb, use_res = case
    Cin = C1 + C2
    dtype = torch.bfloat16
    x = q(rnd((N, Cin, D, H, W), 1), dtype)
    w = q(rnd((Cout, Cin, k, k, k), 2, 1.0 / math.sqrt(Cin * k ** 3)), dtype)
    wp = ops.repack_conv_weight(w.to(DEV), ops.round_up(Cout, 16), Cin, dtype)
    b = rnd((Cout,), 3, 0.1).to(DEV)
    xa = to_act(x[:, :C1], dtype)
    x2a = to_act(x[:, C1:], dtype) if C2 else None
    y = ops.conv3d(xa, wp, b, Cout, k, s, pl, phi, x2=x2a, want_stats=True)
    if not hasattr(y, "cstats"):
        pytest.skip("planner chose split-K / batch-in-tile for this shape: statistics fall back to the stats kernel")
    out = from_act(y, Cout).double()                                   # values as stored
    yy = out.reshape(N, Cout // 8, 8, -1)
    want = torch.stack([yy.sum((2, 3)), (yy * yy).sum((2, 3))], dim=-1).reshape(-1)
    assert relerr(y.cstats, want) <= 1e-5, relerr(y.cstats, want)
    # GroupNorm over cat(y, y2) with 32 groups: group size (Cout + 64)/32 need not divide Cout
    y2 = ops.conv3d(xa, ops.repack_conv_weight(w[:64].contiguous().to(DEV), 64, Cin, dtype), b[:64].contiguous(), 64, k, s,
                    pl, phi, x2=x2a, want_stats=True)
    Ct = Cout + 64
    if (Ct // 32) % 8 == 0 and Ct % 32 == 0 and hasattr(y2, "cstats"):
        gamma, beta = (1.0 + 0.1 * rnd((Ct,), 5)).to(DEV), (0.1 * rnd((Ct,), 6)).to(DEV)
        got = ops.groupnorm(y, gamma, beta, 32, 1e-6, True, x2=y2)
        cat = torch.cat([out, from_act(y2, 64).double()], dim=1)
        ref = F.silu(F.group_norm(cat, 32, gamma.double().cpu(), beta.double().cpu(), eps=1e-6))
        assert relerr(from_act(got, Ct), ref) <= 6e-3


@pytest.mark.parametrize("shape", [(2, (6, 8, 4), 128, 256, 192, 64), (1, (9, 11, 7), 64, 128, 64, 0),
                                   (8, (4, 4, 4), 256, 256, 512, 256), (2, (16, 16, 16), 64, 64, 128, 0)],
                         ids=["two_src_skip", "odd", "batch_tile", "big"])
@pytest.mark.parametrize("kernel", [-1, 3], ids=["auto", "w_halo"])
def test_conv3d_fused_skip_conv(shape, kernel):
    """y = conv3(h) + conv1(cat(sx, sx2)) in one launch == the two convs of a ResNet block with a channel change"""
    import ctypes
    from b200ldm import _lib, ops
    lib = _lib.load()
    lib.b200ldm_debug_set_conv_kernel.argtypes = [ctypes.c_int]
    N, sp, C, Cout, S1, S2 = shape
    dtype = torch.bfloat16
    h = q(rnd((N, C) + sp, 91), dtype)
    sx = q(rnd((N, S1 + S2) + sp, 92), dtype)
    w = q(rnd((Cout, C, 3, 3, 3), 93, 1.0 / math.sqrt(C * 27)), dtype)
    ws = q(rnd((Cout, S1 + S2, 1, 1, 1), 94, 1.0 / math.sqrt(S1 + S2)), dtype)
    b = rnd((Cout,), 95, 0.1)
    ref = F.conv3d(h.double(), w.double(), b.double(), padding=1) + F.conv3d(sx.double(), ws.double())
    wp = ops.repack_conv_weight(w.to(DEV), Cout, C, dtype)
    wsp = ops.repack_conv_weight(ws.to(DEV), Cout, S1 + S2, dtype)
    lib.b200ldm_debug_set_conv_kernel(kernel)
    try:
        y = ops.conv3d(to_act(h, dtype), wp, b.to(DEV), Cout, 3, 1, 1, 1,
                       skip=(to_act(sx[:, :S1], dtype), to_act(sx[:, S1:], dtype) if S2 else None, wsp), want_stats=True)
    finally:
        lib.b200ldm_debug_set_conv_kernel(-1)
    e = relerr(from_act(y, Cout), ref)
    assert e <= 1e-2, e


@pytest.mark.parametrize("case", [(2, (5, 6, 7), 64, 1), (1, (9, 11, 7), 128, 8), (2, (8, 8, 8), 256, 16), (1, (4, 5, 3), 64, 3)],
                         ids=lambda c: f"N{c[0]}_{'x'.join(map(str, c[1]))}_C{c[2]}->{c[3]}")
def test_narrow_conv_tap_expanded_matches_direct_and_torch(case):
    """terminal 3x3x3 convs with <= 16 outputs (UNet out.2, decoder conv_out): 1x1x1 GEMM over stacked taps + tap_gather
    must equal F.conv3d (borders = zero padding) and the direct tensor-core conv"""
    from b200ldm import nets
    N, sp, cin, cout = case
    conv = nets.Conv(cin, cout, terminal=True).to(DEV)
    with torch.no_grad():
        conv.weight.copy_(q(rnd(conv.weight.shape, 41) * (cin * 27) ** -0.5, torch.bfloat16))
        conv.bias.copy_(0.1 * rnd(conv.bias.shape, 42))
    x = q(rnd((N, cin) + sp, 43), torch.bfloat16)
    ref = F.conv3d(x.double(), conv.weight.detach().cpu().double(), conv.bias.detach().cpu().double(), padding=1)
    xa = to_act(x, torch.bfloat16)
    assert nets.TAP_EXPAND
    y = conv.run(xa)
    nets.TAP_EXPAND = False
    try:
        y_direct = conv.run(xa)
    finally:
        nets.TAP_EXPAND = True
    assert y.shape == y_direct.shape == (N,) + sp + (16,)
    assert relerr(from_act(y, cout), ref) <= 1e-2
    assert relerr(from_act(y_direct, cout), ref) <= 1e-2
    if cout < 16:
        assert y[..., cout:].abs().max().item() == 0.0      # padded output channels stay exactly zero


@pytest.mark.parametrize("use_noise", [True, False], ids=["noise_stack", "philox"])
def test_tap_gather_fused_scheduler_update_equals_two_kernels(use_noise):
    """scheduler update inside the gather kernel == gather, then b200ldm_scheduler_step (same coefficient row, same
    noise stream and element indexing); fp32 buffers so the only difference is the order of two roundings"""
    import b200ldm
    from b200ldm import ops
    N, D, H, W, cp, Cl = 2, 3, 4, 5, 4, 4
    z = rnd((N, D, H, W, 27 * cp + 20), 91).to(DEV)            # ldz = 128 > 27*cp
    bias = rnd((16,), 92).to(DEV)
    sched = gc.make_scheduler(b200ldm, "ddpm", 10)
    coef = sched.coef_table(DEV)
    counter = torch.full((1,), 3, dtype=torch.int32, device=DEV)
    x0 = rnd((N, D, H, W, Cl), 93).to(DEV)
    nz = rnd((10 * N, D, H, W, Cl), 94).to(DEV) if use_noise else None
    stride = 0 if nz is None else nz.numel() // 10
    xm_a = torch.zeros((N, D, H, W, 64), device=DEV)
    xm_b = torch.zeros_like(xm_a)
    # two kernels
    xa = x0.clone()
    eps = ops.tap_gather(z, bias, cp, 16)
    ops.scheduler_step(eps, xa, coef, counter, Cl, noise=nz, noise_step_stride=stride, x_model=xm_a, seed=77)
    # one kernel
    xb = x0.clone()
    sp = ops.make_sched_params(xb, coef, counter, Cl, torch.float32, noise=nz, noise_step_stride=stride, x_model=xm_b,
                               seed=77)
    assert ops.tap_gather(z, bias, cp, 16, sched=sp, want_y=False) is None
    assert torch.allclose(xa, xb, atol=1e-5, rtol=1e-5) and torch.allclose(xm_a, xm_b, atol=1e-5, rtol=1e-5)
    assert not torch.equal(xb, x0)
    eps2 = ops.tap_gather(z, bias, cp, 16, sched=ops.make_sched_params(xb.clone(), coef, counter, Cl, torch.float32, seed=1))
    assert torch.equal(eps2, eps)                               # the prediction is still available when asked for


def test_tap_gather_fp32_exact():
    """the gather alone, fp32: y = bias + sum over in-bounds taps, checked against an index-by-index torch restatement"""
    from b200ldm import ops
    N, D, H, W, cp = 2, 3, 4, 5, 2
    z = rnd((N, D, H, W, 64), 44)
    b = rnd((16,), 45)
    y = ops.tap_gather(z.to(DEV), b.to(DEV), cp, 16).cpu()
    zp = F.pad(z.permute(0, 4, 1, 2, 3), (1, 1, 1, 1, 1, 1))          # [N, 64, D+2, H+2, W+2]
    want = torch.zeros(N, D, H, W, 16)
    want[..., :cp] = b[:cp]
    for kd in range(3):
        for kh in range(3):
            for kw in range(3):
                t = (kd * 3 + kh) * 3 + kw
                want[..., :cp] += zp[:, t * cp:(t + 1) * cp, kd:kd + D, kh:kh + H, kw:kw + W].permute(0, 2, 3, 4, 1)
    assert torch.allclose(y, want, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("kernel", [0, 1, 2, 3, 4, 5], ids=["tile_per_cta", "cta_pairs", "persistent", "w_halo", "persistent_wide_epi", "stream_k"])
@pytest.mark.parametrize("case", [(2, (9, 11, 7), 128, 0, 256, 3, 1), (1, (8, 8, 8), 64, 64, 64, 3, 1), (2, (6, 10, 6), 64, 0, 128, 3, 2),
                                  (3, (5, 5, 5), 256, 0, 512, 1, 1), (1, (18, 22, 14), 64, 0, 64, 3, 1),
                                  (8, (4, 4, 4), 128, 0, 192, 3, 1), (5, (3, 4, 2), 64, 64, 320, 3, 1),
                                  (2, (16, 16, 16), 64, 0, 64, 3, 1), (1, (9, 11, 13), 128, 0, 128, 3, 1),
                                  (3, (4, 20, 24), 128, 64, 64, 3, 1), (2, (1, 1, 300), 256, 0, 768, 1, 1),
                                  (1, (6, 6, 6), 128, 128, 512, 3, 1)],
                         ids=lambda c: f"N{c[0]}_{'x'.join(map(str, c[1]))}_C{c[2]}+{c[3]}->{c[4]}_k{c[5]}s{c[6]}")
def test_conv3d_every_tensor_core_kernel(case, kernel):
    """the tcgen05 conv kernels (one CTA per tile, persistent CTA pairs, persistent single CTAs, persistent W-halo) pinned
    one by one (a kernel that does not take a problem leaves it to kernel 0): same tensor, residual + time-embedding +
    GroupNorm chunk statistics included"""
    import ctypes
    from b200ldm import _lib, ops
    lib = _lib.load()
    lib.b200ldm_debug_set_conv_kernel.argtypes = [ctypes.c_int]
    N, sp, C1, C2, Cout, k, stride = case
    C = C1 + C2
    x = q(rnd((N, C) + sp, 81), torch.bfloat16)
    w = q(rnd((Cout, C, k, k, k), 82) * (C * k ** 3) ** -0.5, torch.bfloat16)
    b, temb = 0.1 * rnd((Cout,), 83), 0.1 * rnd((N, Cout), 84)
    pad = k // 2
    ref = F.conv3d(x.double(), w.double(), b.double(), stride=stride, padding=pad) + temb.double()[:, :, None, None, None]
    res = q(rnd(tuple(ref.shape), 85), torch.bfloat16) if stride == 1 else None
    if res is not None:
        ref = ref + res.double()
    xa = to_act(x[:, :C1], torch.bfloat16)
    x2a = to_act(x[:, C1:], torch.bfloat16) if C2 else None
    wp = ops.repack_conv_weight(w.to(DEV), Cout, C, torch.bfloat16, C1 if C2 else None, C1 if C2 else None)
    lib.b200ldm_debug_set_conv_kernel(kernel)
    try:
        y = ops.conv3d(xa, wp, b.to(DEV), Cout, ksize=k, stride=stride, pad_lo=pad, x2=x2a, temb=temb.to(DEV),
                       temb_stride=Cout, residual=None if res is None else to_act(res, torch.bfloat16), want_stats=True)
    finally:
        lib.b200ldm_debug_set_conv_kernel(-1)
    assert relerr(from_act(y, Cout), ref) <= 1e-2
    cs = getattr(y, "cstats", None)
    if cs is not None:      # (sum, sumsq) per 8-channel chunk of the tensor as stored
        yf = from_act(y, Cout).double()
        want = torch.stack([yf.reshape(N, Cout // 8, 8, -1).sum(dim=(2, 3)), (yf ** 2).reshape(N, Cout // 8, 8, -1).sum(dim=(2, 3))], dim=-1)
        assert relerr(cs.reshape(N, Cout // 8, 2).cpu(), want) <= 1e-4


@pytest.mark.skipif(os.environ.get("B200LDM_CONV_AUTOTUNE", "1") == "0", reason="tuning disabled by B200LDM_CONV_AUTOTUNE=0")
def test_conv3d_tuning_is_an_explicit_entry_point():
    """ABI contract (SURVEY.md 8b threading row): b200ldm_conv3d never measures or synchronises -
    it is capture-safe even for a problem signature nothing has been measured for; measuring is b200ldm_conv3d_tune,
    which refuses a capturing stream, and after which the same signature is no longer pending."""
    import ctypes as C
    from b200ldm import _lib, ops
    lib = _lib.load()
    N, sp, Cin, Cout = 3, (5, 6, 7), 128, 192            # a signature no other test uses (fresh in this process)
    x = to_act(q(rnd((N, Cin) + sp, 91), torch.bfloat16), torch.bfloat16)
    w = q(rnd((Cout, Cin, 3, 3, 3), 92) * (Cin * 27) ** -0.5, torch.bfloat16)
    wp = ops.repack_conv_weight(w.to(DEV), Cout, Cin, torch.bfloat16)
    y = torch.empty((N,) + sp + (Cout,), dtype=torch.bfloat16, device=DEV)

    def params():
        p = _lib.ConvParams(x=x.data_ptr(), w=wp.data_ptr(), y=y.data_ptr(), N=N, D=sp[0], H=sp[1], W=sp[2], C1=Cin, C2=0,
                            Cout=Cout, Do=sp[0], Ho=sp[1], Wo=sp[2], ksize=3, stride=1, pad_lo=1, dtype=_lib.BF16,
                            engine=_lib.ENGINE_AUTO)
        nbytes = lib.b200ldm_conv3d_workspace_bytes(C.byref(p))
        ws = torch.empty((max(nbytes, 16),), dtype=torch.uint8, device=DEV)
        p.workspace, p.workspace_bytes = ws.data_ptr(), nbytes
        return p, ws

    p, ws = params()
    assert lib.b200ldm_conv3d_tune_pending(C.byref(p)) == 1
    # 1. the plain entry point inside a stream capture, nothing measured yet: must succeed (cost-model plan, no sync)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        _lib.check(lib.b200ldm_conv3d(C.byref(p), torch.cuda.current_stream().cuda_stream), "conv3d")
        # ... and the measuring entry point must refuse the capturing stream instead of synchronising it
        assert lib.b200ldm_conv3d_tune(C.byref(p), torch.cuda.current_stream().cuda_stream) == 5   # BAD_ARG
    assert lib.b200ldm_conv3d_tune_pending(C.byref(p)) == 1          # calling conv3d did not measure anything
    g.replay()
    torch.cuda.synchronize()
    ref = F.conv3d(ops.to_ncdhw(x, Cin).double().cpu(), w.double(), padding=1)
    assert relerr(from_act(y, Cout), ref) <= 1e-2
    # 2. explicit tuning: result valid, signature settled, a second call is a no-op
    y.zero_()
    _lib.check(lib.b200ldm_conv3d_tune(C.byref(p), torch.cuda.current_stream().cuda_stream), "conv3d_tune")
    torch.cuda.synchronize()
    assert relerr(from_act(y, Cout), ref) <= 1e-2
    assert lib.b200ldm_conv3d_tune_pending(C.byref(p)) == 0
    before = ops.launch_count()
    _lib.check(lib.b200ldm_conv3d_tune(C.byref(p), torch.cuda.current_stream().cuda_stream), "conv3d_tune")
    assert ops.launch_count() == before
    # 3. the remembered choice runs through the plain entry point (workspace re-asked: it may have shrunk)
    p2, ws2 = params()
    y.zero_()
    _lib.check(lib.b200ldm_conv3d(C.byref(p2), torch.cuda.current_stream().cuda_stream), "conv3d")
    torch.cuda.synchronize()
    assert relerr(from_act(y, Cout), ref) <= 1e-2
    # 4. ops.conv3d tunes a fresh signature on its first eager call
    x4 = to_act(q(rnd((2, Cin) + (3, 5, 9), 93), torch.bfloat16), torch.bfloat16)
    p4 = _lib.ConvParams(x=x4.data_ptr(), w=wp.data_ptr(), y=y.data_ptr(), N=2, D=3, H=5, W=9, C1=Cin, C2=0, Cout=Cout,
                         Do=3, Ho=5, Wo=9, ksize=3, stride=1, pad_lo=1, dtype=_lib.BF16, engine=_lib.ENGINE_AUTO)
    assert lib.b200ldm_conv3d_tune_pending(C.byref(p4)) == 1
    ops.conv3d(x4, wp, None, Cout)
    assert lib.b200ldm_conv3d_tune_pending(C.byref(p4)) == 0


@pytest.mark.parametrize("case", [(8, (5, 5, 5), 1024, 0, 1024, 3, True, True, 0), (8, (5, 5, 5), 1024, 1024, 1024, 3, True, False, 1024),
                                  (4, (4, 4, 4), 512, 0, 512, 1, False, True, 512), (2, (6, 6, 6), 256, 0, 512, 3, False, False, 0)],
                         ids=["res_conv", "concat_in_concat_gn", "k1_proj_concat_gn", "216vox"])
def test_split_k_finalize_fused_into_groupnorm(case):
    """The small UNet levels (4^3 / 5^3) are chains of launch-bound kernels: a split-K conv asked to `defer` leaves its
    fp32 slabs to the GroupNorm that follows (b200ldm_groupnorm_small_slabs: finalize + statistics + normalise in one
    launch).  Same tensors as conv -> splitk_finalize -> GroupNorm: the un-normalised output bit for bit, the normalised
    one to bf16 round-off; any other consumer of a pending tensor triggers the standalone finalize."""
    from b200ldm import ops
    N, sp, C1, C2, Cout, k, use_temb, use_res, Cskip = case
    C = C1 + C2
    x = q(rnd((N, C) + sp, 61), torch.bfloat16)
    w = q(rnd((Cout, C, k, k, k), 62) * (C * k ** 3) ** -0.5, torch.bfloat16)
    b = 0.1 * rnd((Cout,), 63)
    temb = 0.3 * rnd((N, Cout), 64) if use_temb else None
    res = q(rnd((N, Cout) + sp, 65), torch.bfloat16) if use_res else None
    skip = q(rnd((N, Cskip) + sp, 66), torch.bfloat16) if Cskip else None     # second GroupNorm source (skip concat)
    gamma, beta = (1 + 0.1 * rnd((Cout + Cskip,), 67)).to(DEV), (0.1 * rnd((Cout + Cskip,), 68)).to(DEV)
    xa = to_act(x[:, :C1], torch.bfloat16)
    x2a = to_act(x[:, C1:], torch.bfloat16) if C2 else None
    wp = ops.repack_conv_weight(w.to(DEV), Cout, C, torch.bfloat16, C1 if C2 else None, C1 if C2 else None)
    kw = dict(ksize=k, stride=1, pad_lo=k // 2, x2=x2a, temb=None if temb is None else temb.to(DEV),
              temb_stride=Cout if use_temb else 0, residual=None if res is None else to_act(res, torch.bfloat16))
    ska = None if skip is None else to_act(skip, torch.bfloat16)

    y0 = ops.conv3d(xa, wp, b.to(DEV), Cout, **kw)                      # settles the measured plan of this signature
    assert getattr(y0, "pending", None) is None
    n0 = ops.groupnorm(y0, gamma, beta, 32, 1e-6, True, x2=ska)
    l0 = ops.launch_count()
    y1 = ops.conv3d(xa, wp, b.to(DEV), Cout, defer=True, **kw)
    pend = getattr(y1, "pending", None)
    if pend is None:
        pytest.skip("the measured plan of this shape does not split K on this GPU")
    n1 = ops.groupnorm(y1, gamma, beta, 32, 1e-6, True, x2=ska)
    assert ops.launch_count() - l0 == 2 and y1.pending is None        # conv + ONE GroupNorm launch, no finalize launch
    torch.cuda.synchronize()
    assert torch.equal(y1, y0)                                          # the raw tensor was written by the fused kernel
    assert relerr(n1.float(), n0.float()) <= 1e-2
    # reference math on top of the raw tensor
    yc = torch.cat([from_act(y0, Cout)] + ([skip] if skip is not None else []), dim=1).double()
    ref = F.silu(F.group_norm(yc, 32, gamma.double().cpu(), beta.double().cpu(), 1e-6))
    assert relerr(from_act(n1, Cout + Cskip), ref) <= 1e-2
    # any other consumer finalises first
    y2 = ops.conv3d(xa, wp, b.to(DEV), Cout, defer=True, **kw)
    assert y2.pending is not None
    assert torch.equal(ops.to_ncdhw(y2, Cout), ops.to_ncdhw(y0, Cout)) and y2.pending is None
    y3 = ops.conv3d(xa, wp, b.to(DEV), Cout, defer=True, **kw)          # as the x2 source of a GroupNorm
    n3 = ops.groupnorm(y0, (1 + 0.1 * rnd((2 * Cout,), 69)).to(DEV), torch.zeros(2 * Cout, device=DEV), 32, 1e-6, False, x2=y3)
    n4 = ops.groupnorm(y0, (1 + 0.1 * rnd((2 * Cout,), 69)).to(DEV), torch.zeros(2 * Cout, device=DEV), 32, 1e-6, False, x2=y0)
    assert y3.pending is None and torch.equal(n3, n4)


@pytest.mark.parametrize("splits", [2, 3, 8])
@pytest.mark.parametrize("slab_case", [(8, (5, 5, 5), 256, 0, 512, 3), (2, (10, 10, 10), 128, 64, 256, 3), (8, (4, 4, 4), 128, 0, 192, 3),
                                       (3, (3, 5, 2), 256, 0, 64, 1), (5, (4, 4, 4), 64, 64, 320, 3)],
                         ids=lambda c: f"N{c[0]}_{'x'.join(map(str, c[1]))}_C{c[2]}+{c[3]}->{c[4]}_k{c[5]}")
def test_split_k_slabs_by_tma_and_finalize_with_statistics(slab_case, splits):
    """Split-K forced on the one-CTA-per-tile kernel: the fp32 partial sums leave through smem + TMA stores (one map over
    all slabs; boxes with batch folded in - 4^3, N % tn != 0 falls back to direct stores), the finalize pass adds bias /
    time embedding / residual AND accumulates the GroupNorm chunk sums of the bf16 tensor it writes."""
    import ctypes
    from b200ldm import _lib, ops
    lib = _lib.load()
    lib.b200ldm_debug_set_conv_kernel.argtypes = [ctypes.c_int]
    lib.b200ldm_debug_set_conv_splits.argtypes = [ctypes.c_int]
    N, sp, C1, C2, Cout, k = slab_case
    C = C1 + C2
    x = q(rnd((N, C) + sp, 71), torch.bfloat16)
    w = q(rnd((Cout, C, k, k, k), 72) * (C * k ** 3) ** -0.5, torch.bfloat16)
    b, temb = 0.1 * rnd((Cout,), 73), 0.1 * rnd((N, Cout), 74)
    ref = F.conv3d(x.double(), w.double(), b.double(), padding=k // 2) + temb.double()[:, :, None, None, None]
    res = q(rnd(tuple(ref.shape), 75), torch.bfloat16)
    ref = ref + res.double()
    xa = to_act(x[:, :C1], torch.bfloat16)
    x2a = to_act(x[:, C1:], torch.bfloat16) if C2 else None
    wp = ops.repack_conv_weight(w.to(DEV), Cout, C, torch.bfloat16, C1 if C2 else None, C1 if C2 else None)
    lib.b200ldm_debug_set_conv_kernel(0)
    lib.b200ldm_debug_set_conv_splits(splits)
    try:
        y = ops.conv3d(xa, wp, b.to(DEV), Cout, ksize=k, stride=1, pad_lo=k // 2, x2=x2a, temb=temb.to(DEV),
                       temb_stride=Cout, residual=to_act(res, torch.bfloat16), want_stats=True)
    finally:
        lib.b200ldm_debug_set_conv_kernel(-1)
        lib.b200ldm_debug_set_conv_splits(0)
    assert relerr(from_act(y, Cout), ref) <= 1e-2
    cs = getattr(y, "cstats", None)
    yf = from_act(y, Cout).double()
    if (C // 64) * k ** 3 >= splits:               # the factor was applicable (at least one k-iteration per slab)
        assert cs is not None                      # accumulated by the finalize pass of the same b200ldm_conv3d call
    if cs is not None:
        want = torch.stack([yf.reshape(N, Cout // 8, 8, -1).sum(dim=(2, 3)), (yf ** 2).reshape(N, Cout // 8, 8, -1).sum(dim=(2, 3))], dim=-1)
        assert relerr(cs.reshape(N, Cout // 8, 2).cpu(), want) <= 1e-4
    # the GroupNorm that follows normalises from those sums alone
    gamma, beta = (1 + 0.1 * rnd((Cout,), 76)).to(DEV), (0.1 * rnd((Cout,), 77)).to(DEV)
    got = ops.groupnorm(y, gamma, beta, 32, 1e-6, True)
    gref = F.silu(F.group_norm(yf, 32, gamma.double().cpu(), beta.double().cpu(), 1e-6))
    assert relerr(from_act(got, Cout), gref) <= 1e-2


@pytest.mark.parametrize("sk_case", [(8, (10, 10, 10), 256, 0, 256, 3, 1, 0), (8, (20, 20, 20), 64, 0, 256, 3, 1, 0),
                                     (3, (9, 11, 7), 128, 128, 512, 3, 1, 128), (2, (12, 10, 8), 128, 0, 256, 3, 2, 0),
                                     (1, (1, 1, 1000), 256, 0, 768, 1, 1, 0), (8, (5, 5, 5), 512, 0, 256, 3, 1, 0)],
                         ids=lambda c: f"N{c[0]}_{'x'.join(map(str, c[1]))}_C{c[2]}+{c[3]}->{c[4]}_k{c[5]}s{c[6]}_skip{c[7]}")
def test_conv3d_stream_k(sk_case):
    """conv_sk.cu pinned: every CTA runs an equal share of ALL k-iterations, tiles cut by a span boundary are completed
    through partial tiles in the workspace (contributor -> owner inside the launch).  Shapes with several CTAs per tile,
    whole tiles between two partial ones, two sources + fused skip conv, stride 2, a 1x1x1 projection; bias + time
    embedding + residual + GroupNorm chunk sums on the owner's epilogue; repeated launches reuse the (self-cleaning)
    flags; bitwise repeatable."""
    import ctypes
    from b200ldm import _lib, ops
    lib = _lib.load()
    lib.b200ldm_debug_set_conv_kernel.argtypes = [ctypes.c_int]
    N, sp, C1, C2, Cout, k, stride, Cs = sk_case
    C = C1 + C2
    x = q(rnd((N, C) + sp, 31), torch.bfloat16)
    w = q(rnd((Cout, C, k, k, k), 32) * (C * k ** 3) ** -0.5, torch.bfloat16)
    b, temb = 0.1 * rnd((Cout,), 33), 0.1 * rnd((N, Cout), 34)
    pad = k // 2
    ref = F.conv3d(x.double(), w.double(), b.double(), stride=stride, padding=pad) + temb.double()[:, :, None, None, None]
    res = q(rnd(tuple(ref.shape), 35), torch.bfloat16) if stride == 1 else None
    if res is not None:
        ref = ref + res.double()
    skip = None
    if Cs:
        sx = q(rnd((N, Cs) + sp, 36), torch.bfloat16)
        sw = q(rnd((Cout, Cs, 1, 1, 1), 37) * Cs ** -0.5, torch.bfloat16)
        ref = ref + F.conv3d(sx.double(), sw.double())
        skip = (to_act(sx, torch.bfloat16), None, ops.repack_conv_weight(sw.to(DEV), Cout, Cs, torch.bfloat16))
    xa = to_act(x[:, :C1], torch.bfloat16)
    x2a = to_act(x[:, C1:], torch.bfloat16) if C2 else None
    wp = ops.repack_conv_weight(w.to(DEV), Cout, C, torch.bfloat16, C1 if C2 else None, C1 if C2 else None)
    kw = dict(ksize=k, stride=stride, pad_lo=pad, x2=x2a, temb=temb.to(DEV), temb_stride=Cout,
              residual=None if res is None else to_act(res, torch.bfloat16), want_stats=True, skip=skip)
    lib.b200ldm_debug_set_conv_kernel(5)
    try:
        ys = [ops.conv3d(xa, wp, b.to(DEV), Cout, **kw) for _ in range(3)]
        torch.cuda.synchronize()
    finally:
        lib.b200ldm_debug_set_conv_kernel(-1)
    y = ys[0]
    assert relerr(from_act(y, Cout), ref) <= 1e-2
    assert torch.equal(ys[1], y) and torch.equal(ys[2], y)
    cs = getattr(y, "cstats", None)
    if cs is not None:
        yf = from_act(y, Cout).double()
        want = torch.stack([yf.reshape(N, Cout // 8, 8, -1).sum(dim=(2, 3)), (yf ** 2).reshape(N, Cout // 8, 8, -1).sum(dim=(2, 3))], dim=-1)
        assert relerr(cs.reshape(N, Cout // 8, 2).cpu(), want) <= 1e-4
    # against the one-CTA-per-tile kernel: same products, different summation order of the fp32 partial tiles
    lib.b200ldm_debug_set_conv_kernel(0)
    try:
        y0 = ops.conv3d(xa, wp, b.to(DEV), Cout, **kw)
    finally:
        lib.b200ldm_debug_set_conv_kernel(-1)
    assert relerr(y.float(), y0.float()) <= 8e-3


@pytest.mark.parametrize("case", [(2, (9, 11, 7), 1, 64, 8), (1, (16, 12, 20), 2, 64, 64), (3, (4, 5, 6), 1, 128, 16)],
                         ids=lambda c: f"N{c[0]}_{'x'.join(map(str, c[1]))}_cin{c[2]}->{c[3]}_Cp{c[4]}")
def test_narrow_input_conv_as_packed_taps(case):
    """AutoencoderKL conv_in (1 or 2 input channels): ops.pack_taps puts the 27 taps of a voxel's neighbourhood into one
    64-channel row (zero padding at the borders), the conv runs as ONE 1x1x1 k-iteration with tap-major weights - same
    products as the 27-tap conv over the zero-padded input, against fp64 F.conv3d and against the regular kernel."""
    from b200ldm import nets, ops
    N, sp, cin, cout, Cp = case
    x = q(rnd((N, cin) + sp, 41), torch.bfloat16)
    conv = nets.Conv(cin, cout).to(DEV)
    with torch.no_grad():
        conv.weight.copy_(q(rnd((cout, cin, 3, 3, 3), 42) * (cin * 27) ** -0.5, torch.bfloat16))
        conv.bias.copy_(0.1 * rnd((cout,), 43))
    ref = F.conv3d(x.double(), conv.weight.detach().double().cpu(), conv.bias.detach().double().cpu(), padding=1)
    xa = ops.to_ndhwc(x.to(DEV), Cp, torch.bfloat16)
    packed = ops.pack_taps(xa, cin)
    # the packed tensor itself: channel t * cin + c = x at tap t
    xp = F.pad(x, (1, 1, 1, 1, 1, 1))
    D, H, W = sp
    want = torch.zeros((N, D, H, W, 64))
    for t in range(27):
        kd, kh, kw = t // 9, (t // 3) % 3, t % 3
        want[..., t * cin:(t + 1) * cin] = xp[:, :, kd:kd + D, kh:kh + H, kw:kw + W].permute(0, 2, 3, 4, 1)
    assert torch.equal(packed.float().cpu(), want)
    assert conv.tap_packed(xa)
    y = conv.run(xa, want_stats=True)
    assert relerr(from_act(y, cout), ref) <= 1e-2
    keep = nets.TAP_PACK
    nets.TAP_PACK = False
    try:
        y0 = conv.run(ops.to_ndhwc(x.to(DEV), 64, torch.bfloat16))
    finally:
        nets.TAP_PACK = keep
    assert relerr(y.float(), y0.float()) <= 8e-3


CAND_CASES = [(2, (6, 8, 4), 64, 0, 64, 3, 1, 0), (2, (3, 4, 2), 128, 0, 128, 3, 1, 0), (2, (3, 4, 2), 128, 128, 128, 3, 1, 256),
              (2, (6, 8, 4), 64, 128, 64, 3, 1, 192), (2, (6, 8, 4), 64, 0, 64, 3, 2, 0), (2, (1, 1, 24), 128, 0, 384, 1, 1, 0),
              (8, (5, 5, 5), 256, 0, 512, 3, 1, 0), (3, (10, 10, 10), 128, 0, 256, 3, 1, 0), (1, (9, 11, 7), 256, 0, 256, 1, 1, 0)]


@pytest.mark.parametrize("case", CAND_CASES, ids=lambda c: f"N{c[0]}_{'x'.join(map(str, c[1]))}_C{c[2]}+{c[3]}->{c[4]}_k{c[5]}s{c[6]}_skip{c[7]}")
def test_conv3d_every_tuner_candidate(case):
    """b200ldm_conv3d_tune picks among candidates by TIME; every one of them must be RIGHT.  Each choice code the tuner
    can build (kernel 0 / 2 / 3 / 4 / 5, split-K factor x tile height, narrower Cout tile x tile height / persistent) is
    forced in turn on shapes of the tiny and the real UNets (a candidate that does not apply falls back to the plan it
    would fall back to in the tuner): output vs fp64 F.conv3d, GroupNorm chunk sums vs the stored tensor."""
    import ctypes
    from b200ldm import _lib, ops
    lib = _lib.load()
    lib.b200ldm_debug_set_conv_choice.argtypes = [ctypes.c_int]
    N, sp, C1, C2, Cout, k, stride, Cs = case
    C = C1 + C2
    x = q(rnd((N, C) + sp, 51), torch.bfloat16)
    w = q(rnd((Cout, C, k, k, k), 52) * (C * k ** 3) ** -0.5, torch.bfloat16)
    b, temb = 0.1 * rnd((Cout,), 53), 0.1 * rnd((N, Cout), 54)
    pad = k // 2
    ref = F.conv3d(x.double(), w.double(), b.double(), stride=stride, padding=pad) + temb.double()[:, :, None, None, None]
    res = q(rnd(tuple(ref.shape), 55), torch.bfloat16) if (stride == 1 and not Cs) else None
    if res is not None:
        ref = ref + res.double()
    skip = None
    if Cs:
        sx = q(rnd((N, Cs) + sp, 56), torch.bfloat16)
        sw = q(rnd((Cout, Cs, 1, 1, 1), 57) * Cs ** -0.5, torch.bfloat16)
        ref = ref + F.conv3d(sx.double(), sw.double())
        skip = (to_act(sx, torch.bfloat16), None, ops.repack_conv_weight(sw.to(DEV), Cout, Cs, torch.bfloat16))
    xa = to_act(x[:, :C1], torch.bfloat16)
    x2a = to_act(x[:, C1:], torch.bfloat16) if C2 else None
    wp = ops.repack_conv_weight(w.to(DEV), Cout, C, torch.bfloat16, C1 if C2 else None, C1 if C2 else None)
    kw = dict(ksize=k, stride=stride, pad_lo=pad, x2=x2a, temb=temb.to(DEV), temb_stride=Cout,
              residual=None if res is None else to_act(res, torch.bfloat16), want_stats=True, skip=skip)
    codes = [0, 2, 3, 4, 5]
    codes += [(sp_ << 4) | (mt << 12) for mt in (1, 2) for sp_ in (1, 2, 3, 4, 6, 8, 12, 16)]
    codes += [(bc << 20) | v for bc in (1, 2) for v in (1 << 12, 2 << 12, 2)]
    bad = []
    for code in codes:
        lib.b200ldm_debug_set_conv_choice(code)
        try:
            y = ops.conv3d(xa, wp, b.to(DEV), Cout, **kw)
            torch.cuda.synchronize()
        finally:
            lib.b200ldm_debug_set_conv_choice(-1)
        e = relerr(from_act(y, Cout), ref)
        cs = getattr(y, "cstats", None)
        es = 0.0
        if cs is not None:
            yf = from_act(y, Cout).double()
            want = torch.stack([yf.reshape(N, Cout // 8, 8, -1).sum(dim=(2, 3)), (yf ** 2).reshape(N, Cout // 8, 8, -1).sum(dim=(2, 3))], dim=-1)
            es = relerr(cs.reshape(N, Cout // 8, 2).cpu(), want)
        if not (e <= 1e-2 and es <= 1e-4):
            bad.append((hex(code), float(e), float(es)))
    assert not bad, bad


def test_conv3d_stream_k_under_concurrent_streams():
    """forward progress of the stream-K kernel when several of its launches share the GPU: owners poll the flags of
    contributors that may not be resident yet; CTAs are dispatched in order and a contribution is always posted before its
    CTA waits for anything, so every launch completes (scripts/stress_streamk.py runs 200 rounds) - and bit for bit equal
    to the launch that had the GPU to itself."""
    import ctypes
    from b200ldm import _lib, ops
    lib = _lib.load()
    lib.b200ldm_debug_set_conv_kernel.argtypes = [ctypes.c_int]
    probs = []
    for N, sp, C, Cout in [(8, (10, 10, 10), 256, 256), (4, (20, 20, 20), 256, 256), (8, (5, 5, 5), 512, 512)]:
        x = to_act(q(rnd((N, C) + sp, 61), torch.bfloat16), torch.bfloat16)
        w = ops.repack_conv_weight((q(rnd((Cout, C, 3, 3, 3), 62), torch.bfloat16) * (C * 27) ** -0.5).to(DEV), Cout, C, torch.bfloat16)
        probs.append((x, w, torch.zeros(Cout, device=DEV), Cout))
    lib.b200ldm_debug_set_conv_kernel(5)
    try:
        refs = [ops.conv3d(x, w, b, co) for x, w, b, co in probs]
        torch.cuda.synchronize()
        streams = [torch.cuda.Stream() for _ in range(3)]
        for it in range(30):
            outs = []
            for si, st in enumerate(streams):
                with torch.cuda.stream(st):
                    k = (it + si) % len(probs)
                    x, w, b, co = probs[k]
                    outs.append((k, ops.conv3d(x, w, b, co)))
            torch.cuda.synchronize()
            assert all(torch.equal(y, refs[k]) for k, y in outs)
    finally:
        lib.b200ldm_debug_set_conv_kernel(-1)


def test_conv3d_rejects_bad_arguments():
    from b200ldm import ops
    x = torch.zeros((1, 4, 4, 4, 48), dtype=torch.bfloat16, device=DEV)
    w = torch.zeros((27, 64, 48), dtype=torch.bfloat16, device=DEV)
    with pytest.raises(RuntimeError, match="multiples of 64"):
        ops.conv3d(x, w, None, 64)
    x = torch.zeros((1, 4, 4, 4, 64), dtype=torch.bfloat16, device=DEV)
    with pytest.raises(RuntimeError, match="ksize"):
        ops.conv3d(x, w, None, 64, ksize=5, pad_lo=2)
    with pytest.raises(RuntimeError, match="inconsistent"):
        ops.conv3d(x, w, None, 64, ksize=3, pad_lo=1, pad_hi=3)


# ------------------------------------------------------------------------------------------
# GroupNorm (+SiLU)
# ------------------------------------------------------------------------------------------
GN_CASES = [(2, (4, 6, 5), 64, 0), (1, (8, 8, 8), 256, 0), (2, (3, 5, 7), 512, 256), (1, (4, 4, 4), 1024, 1024),
            (3, (2, 3, 5), 1024, 512), (1, (16, 16, 16), 128, 0)]


@pytest.mark.parametrize("path", ["auto", "two_pass"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("silu", [True, False])
@pytest.mark.parametrize("case", GN_CASES, ids=[f"N{c[0]}_C{c[2]}+{c[3]}" for c in GN_CASES])
def test_groupnorm(case, silu, dtype, path):
    """auto: small tensors take the single-launch kernel (one CTA per (group, sample)); two_pass: stats + apply"""
    from b200ldm import ops
    N, sp, C1, C2 = case
    C = C1 + C2
    x = q(rnd((N, C) + sp, 11) * 2.0 + 0.7, dtype)
    gamma, beta = 1.0 + 0.1 * rnd((C,), 12), 0.1 * rnd((C,), 13)
    ref = F.group_norm(x.double(), 32, gamma.double(), beta.double(), eps=1e-6)
    if silu:
        ref = F.silu(ref)
    xa = to_act(x[:, :C1], dtype)
    x2a = to_act(x[:, C1:], dtype) if C2 else None
    keep = ops.GN_SMALL_MAX_S
    ops.GN_SMALL_MAX_S = keep if path == "auto" else 0
    try:
        y = ops.groupnorm(xa, gamma.to(DEV), beta.to(DEV), 32, 1e-6, silu, x2=x2a)
    finally:
        ops.GN_SMALL_MAX_S = keep
    e = relerr(from_act(y, C), ref)
    assert e <= (1e-5 if dtype == torch.float32 else 6e-3), e


# ------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------
ATT_CASES = [(2, 2, 100, 64), (1, 4, 693, 64), (1, 1, 300, 256), (2, 1, 129, 128), (1, 8, 512, 64), (1, 2, 64, 32)]


def _att_ref(qkv, H, d):
    B, L, _ = qkv.shape
    qq, kk, vv = (t.reshape(B, L, H, d).permute(0, 2, 1, 3).double() for t in qkv.chunk(3, dim=-1))
    att = torch.softmax(qq @ kk.transpose(-1, -2) * d ** -0.5, dim=-1)
    return (att @ vv).permute(0, 2, 1, 3).reshape(B, L, H * d)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("case", ATT_CASES, ids=[f"B{c[0]}H{c[1]}L{c[2]}d{c[3]}" for c in ATT_CASES])
def test_attention_simt(case, dtype):
    from b200ldm import ops
    B, H, L, d = case
    qkv = q(rnd((B, L, 3 * H * d), 21), dtype)
    o = ops.attention(qkv.to(DEV).to(dtype), H, d, d ** -0.5, engine=ops.ENGINE_SIMT)
    e = relerr(o.float(), _att_ref(qkv, H, d))
    assert e <= (2e-5 if dtype == torch.float32 else 6e-3), e


# beyond the SIMT list: BASELINE cfg 2 attention (B8 H4 L5544 d64), AutoencoderKL d = 256 at L = 4096, the d = 128
# ping-pong kernel at L >= 4096 (config_train_multigpu / config_train_stable AutoencoderKL attention),
# head dim 32 (config_train_stable.json:47) on the 64-column tiles: even / odd / many heads, short and long sequences
TC_ATT_CASES = ATT_CASES[:5] + [(8, 4, 5544, 64), (1, 1, 4096, 256), (2, 16, 64, 64), (1, 1, 4096, 128),
                                (2, 1, 4500, 128), (1, 2, 64, 32), (2, 3, 100, 32), (1, 16, 1000, 32), (2, 16, 125, 32),
                                (1, 1, 300, 32)]


@pytest.mark.parametrize("case", TC_ATT_CASES, ids=lambda c: f"B{c[0]}H{c[1]}L{c[2]}d{c[3]}")
def test_attention_tcgen05_bf16(case):
    from b200ldm import ops
    B, H, L, d = case
    qkv = q(rnd((B, L, 3 * H * d), 22), torch.bfloat16)
    dev_qkv = qkv.to(DEV).to(torch.bfloat16)
    o = ops.attention(dev_qkv, H, d, d ** -0.5, engine=ops.ENGINE_TCGEN05)
    if L <= 1024:
        ref = _att_ref(qkv, H, d)
    else:  # big cases: the SIMT engine (validated above against fp64 torch) is the checker
        ref = ops.attention(dev_qkv, H, d, d ** -0.5, engine=ops.ENGINE_SIMT).float()
    e = relerr(o.float(), ref)
    assert e <= 1e-2, e


@pytest.mark.parametrize("factor", [-1, 2, 3, 5, 8])
@pytest.mark.parametrize("case", [(1, 1, 4500, 256), (1, 1, 20000, 256), (2, 3, 1100, 64), (1, 5, 700, 32), (3, 1, 333, 256)],
                         ids=lambda c: f"B{c[0]}H{c[1]}L{c[2]}d{c[3]}")
def test_attention_split_kv_of_the_partial_wave(case, factor):
    """b200ldm_attention_ws: with scratch, the query tiles of the last (partial) wave run as `factor` key ranges each
    (un-normalised O + reference max + row sum per part) and a merge launch combines them.  Forced factors on shapes
    whose tile count is below / above one wave, ragged last query and key tiles, head dims 32 / 64 / 256; factor -1 = the
    planner's own decision (L = 20000, d = 256: 157 tiles on 148 SMs).  Same result as whole tiles to bf16 round-off."""
    import ctypes
    from b200ldm import _lib, ops
    lib = _lib.load()
    lib.b200ldm_debug_set_attn_split.argtypes = [ctypes.c_int]
    B, H, L, d = case
    qkv = q(rnd((B, L, 3 * H * d), 28), torch.bfloat16)
    qkv[:, L // 2:, H * d: 2 * H * d] *= 2.5          # later keys dominate: the parts end with different reference maxima
    dev_qkv = q(qkv, torch.bfloat16).to(DEV).to(torch.bfloat16)
    lib.b200ldm_debug_set_attn_split(0)
    try:
        whole = ops.attention(dev_qkv, H, d, d ** -0.5, engine=ops.ENGINE_TCGEN05)
        lib.b200ldm_debug_set_attn_split(factor)
        ws = lib.b200ldm_attention_workspace_bytes(B, H, L, L, d, _lib.BF16, ops.ENGINE_TCGEN05)
        l0 = ops.launch_count()
        split = ops.attention(dev_qkv, H, d, d ** -0.5, engine=ops.ENGINE_TCGEN05)
        launches = ops.launch_count() - l0
    finally:
        lib.b200ldm_debug_set_attn_split(-1)
    bn = 64 if (d == 256 or L > 128) else 128
    if (factor >= 2 and -(-L // bn) >= factor) or case == (1, 1, 20000, 256):
        assert ws > 0, ws
    assert launches == (2 if ws > 0 else 1), (ws, launches)     # attention (+ merge)
    assert relerr(split.float(), whole.float()) <= 6e-3
    if L <= 1200:
        assert relerr(split.float(), _att_ref(q(qkv, torch.bfloat16), H, d)) <= 1e-2


def test_attention_auto_engine_never_falls_back_to_simt():
    """north_star: no silent CUDA-core fallback on the bf16 product path.  ENGINE_AUTO with a head
    dim that has no tcgen05 instance raises; head dim 32 runs on tcgen05 and agrees with the explicit engine."""
    from b200ldm import ops
    qkv = q(rnd((1, 70, 3 * 512), 26), torch.bfloat16).to(DEV).to(torch.bfloat16)
    with pytest.raises(RuntimeError, match="BAD_SHAPE"):
        ops.attention(qkv, 1, 512, 512 ** -0.5)
    ops.attention(qkv, 1, 512, 512 ** -0.5, engine=ops.ENGINE_SIMT)      # explicit validation engine still works
    qkv32 = q(rnd((2, 200, 3 * 4 * 32), 27), torch.bfloat16).to(DEV).to(torch.bfloat16)
    a = ops.attention(qkv32, 4, 32, 32 ** -0.5)
    b = ops.attention(qkv32, 4, 32, 32 ** -0.5, engine=ops.ENGINE_TCGEN05)
    assert torch.equal(a, b)


def test_attention_tcgen05_peaked_softmax():
    """large logits: exercises the lazy-rescale path (running max jumps by > 2^8 between KV tiles)"""
    from b200ldm import ops
    B, H, L, d = 1, 2, 700, 64
    qkv = rnd((B, L, 3 * H * d), 23)
    qkv[:, :, : 2 * H * d] *= 6.0
    qkv[:, 400:, H * d: 2 * H * d] *= 3.0   # later keys dominate
    qkv = q(qkv, torch.bfloat16)
    o = ops.attention(qkv.to(DEV).to(torch.bfloat16), H, d, d ** -0.5, engine=ops.ENGINE_TCGEN05)
    e = relerr(o.float(), _att_ref(qkv, H, d))
    assert e <= 1e-2, e


def _xatt_ref(qm, kv, H, d):
    B, Lq, _ = qm.shape
    Lk = kv.shape[1]
    qq = qm.double().view(B, Lq, H, d).permute(0, 2, 1, 3)
    kk, vv = (kv.double()[..., i * H * d:(i + 1) * H * d].reshape(B, Lk, H, d).permute(0, 2, 1, 3) for i in range(2))
    att = torch.softmax(qq @ kk.transpose(-1, -2) * d ** -0.5, dim=-1)
    return (att @ vv).permute(0, 2, 1, 3).reshape(B, Lq, H * d)


# B, H, Lq, Lk, d: context of one token, a ragged single tile, several 64-key tiles (d = 64 two-CTA kernel), head dims
# 128 / 256, more keys than queries
XATT_CASES = [(2, 2, 300, 1, 64), (2, 2, 384, 7, 64), (1, 4, 130, 77, 64), (2, 2, 200, 333, 64), (1, 2, 257, 50, 128),
              (1, 2, 300, 400, 128), (1, 1, 190, 100, 256), (2, 2, 5, 9, 64), (2, 4, 150, 77, 32), (1, 3, 129, 200, 32)]


@pytest.mark.parametrize("case", XATT_CASES, ids=lambda c: f"B{c[0]}H{c[1]}Lq{c[2]}Lk{c[3]}d{c[4]}")
def test_cross_attention_tcgen05_bf16(case):
    from b200ldm import ops
    B, H, Lq, Lk, d = case
    qm, kv = q(rnd((B, Lq, H * d), 24), torch.bfloat16), q(rnd((B, Lk, 2 * H * d), 25), torch.bfloat16)
    o = ops.cross_attention(qm.to(DEV).bfloat16(), kv.to(DEV).bfloat16(), H, d, d ** -0.5, engine=ops.ENGINE_TCGEN05)
    e = relerr(o.float(), _xatt_ref(qm, kv, H, d))
    assert e <= 1e-2, e


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("case", [(2, 2, 70, 7, 32), (1, 2, 130, 45, 64), (1, 1, 33, 200, 128)],
                         ids=lambda c: f"B{c[0]}H{c[1]}Lq{c[2]}Lk{c[3]}d{c[4]}")
def test_cross_attention_simt(case, dtype):
    from b200ldm import ops
    B, H, Lq, Lk, d = case
    qm, kv = q(rnd((B, Lq, H * d), 26), dtype), q(rnd((B, Lk, 2 * H * d), 27), dtype)
    o = ops.cross_attention(qm.to(DEV).to(dtype), kv.to(DEV).to(dtype), H, d, d ** -0.5, engine=ops.ENGINE_SIMT)
    e = relerr(o.float(), _xatt_ref(qm, kv, H, d))
    assert e <= (2e-5 if dtype == torch.float32 else 6e-3), e


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("shape", [(5, 64), (1000, 128), (777, 320), (64, 1024), (9, 2048), (3, 4096)],
                         ids=lambda s: f"M{s[0]}C{s[1]}")
def test_layernorm(shape, dtype):
    from b200ldm import ops
    M, Cc = shape
    x = q(rnd((M, Cc), 28) * 3.0 + 1.5, dtype)
    g, b = 1.0 + 0.1 * rnd((Cc,), 29), 0.1 * rnd((Cc,), 30)
    y = ops.layernorm(x.to(DEV).to(dtype), g.to(DEV), b.to(DEV), 1e-5)
    ref = F.layer_norm(x.double(), (Cc,), g.double(), b.double(), 1e-5)
    e = relerr(y.float(), ref)
    assert e <= (2e-6 if dtype == torch.float32 else 5e-3), e


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_geglu(dtype):
    from b200ldm import ops
    h = q(rnd((333, 2 * 512), 31) * 2.0, dtype)
    y = ops.geglu(h.to(DEV).to(dtype))
    ref = h[:, :512].double() * F.gelu(h[:, 512:].double())
    e = relerr(y.float(), ref)
    assert e <= (2e-6 if dtype == torch.float32 else 5e-3), e


def test_embedding_add():
    from b200ldm import ops
    a, table = rnd((3, 256), 32), rnd((5, 256), 33)
    idx = torch.tensor([4, 0, 2])
    got = ops.embedding_add(a.to(DEV), table.to(DEV), idx.to(DEV))
    assert torch.equal(got.cpu(), a + table[idx])


# ------------------------------------------------------------------------------------------
# element-wise family
# ------------------------------------------------------------------------------------------
def test_layout_bridges_roundtrip_and_padding():
    from b200ldm import ops
    x = rnd((2, 5, 3, 4, 6), 31)
    a = ops.to_ndhwc(x.to(DEV), 64, torch.float32)
    assert a.shape == (2, 3, 4, 6, 64)
    assert torch.equal(a[..., :5].cpu(), x.permute(0, 2, 3, 4, 1))
    assert a[..., 5:].abs().max().item() == 0.0
    assert torch.equal(ops.to_ncdhw(a, 5).cpu(), x)
    ab = ops.to_ndhwc(x.to(DEV), 8, torch.bfloat16)
    assert torch.equal(ops.to_ncdhw(ab, 5).cpu(), x.bfloat16().float())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_upsample_nearest2x(dtype):
    from b200ldm import ops
    x = q(rnd((2, 64, 3, 5, 4), 32), dtype)
    y = ops.upsample_nearest2x(to_act(x, dtype))
    assert torch.equal(from_act(y, 64), F.interpolate(x, scale_factor=2.0, mode="nearest"))


def test_timestep_embedding_and_silu():
    from b200ldm import ops
    from oracle.nets import get_timestep_embedding
    ts = torch.tensor([0.0, 1.0, 37.0, 500.0, 999.0])
    got = ops.timestep_embedding(ts.to(DEV), 256).cpu()
    assert (got - get_timestep_embedding(ts, 256)).abs().max().item() <= 2e-4  # libm-level differences on args ~ 1e3
    x = rnd((1000,), 33, 3.0)
    assert torch.allclose(ops.silu_f32(x.to(DEV)).cpu(), F.silu(x), atol=1e-6, rtol=1e-5)


def test_repack_weight_layout():
    from b200ldm import ops
    w = rnd((5, 6, 3, 3, 3), 34)
    p = ops.repack_conv_weight(w.to(DEV), 16, 128, torch.float32, cin_split=4, cin_split_pad=64).cpu()
    assert p.shape == (27, 16, 128)
    want = torch.zeros(27, 16, 128)
    wt = w.reshape(5, 6, 27).permute(2, 0, 1)
    want[:, :5, :4] = wt[:, :, :4]
    want[:, :5, 64:66] = wt[:, :, 4:]
    assert torch.equal(p, want)


@pytest.mark.parametrize("kind", ["ddpm", "ddim"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_scheduler_step_kernel_vs_oracle(kind, dtype):
    import b200ldm
    from b200ldm import ops
    n = 10 if kind == "ddpm" else 50
    ours, ref = gc.make_scheduler(b200ldm, kind, n), gc.make_scheduler(oracle, kind, n)
    coef = ours.coef_table(DEV)
    x, nz = rnd((2, 4, 3, 5, 4), 41), rnd((n, 2, 4, 3, 5, 4), 42)
    eps = q(rnd((2, 4, 3, 5, 4), 43), dtype)
    xa = ops.to_ndhwc(x.to(DEV), 4, torch.float32)
    ea = ops.to_ndhwc(eps.to(DEV), 16, dtype)              # eps buffer wider than C, like the UNet output
    xm = torch.zeros((2, 3, 5, 4, 64), dtype=dtype, device=DEV)
    nza = ops.to_ndhwc(nz.reshape((n * 2, 4, 3, 5, 4)).to(DEV), 4, torch.float32)
    counter = torch.zeros(1, dtype=torch.int32, device=DEV)
    want = x
    for i, t in enumerate(ref.timesteps):
        x0 = torch.empty_like(xa)
        ops.scheduler_step(ea, xa, coef, counter, 4, noise=nza, noise_step_stride=nza.numel() // n, x0_out=x0,
                           x_model=xm)
        ops.counter_add(counter, 1)
        want, want0 = ref.step(eps, t, want, noise=nz[i])
        assert relerr(ops.to_ncdhw(xa, 4), want) <= 2e-5, (i, relerr(ops.to_ncdhw(xa, 4), want))
        assert relerr(ops.to_ncdhw(x0, 4), want0) <= 2e-5
        assert torch.equal(xm[..., :4].float(), xa.to(dtype).float())
        assert xm[..., 4:].abs().max().item() == 0
    assert int(counter.item()) == n


@pytest.mark.parametrize("pred", ["sample", "v_prediction"])
@pytest.mark.parametrize("kind", ["ddpm", "ddim"])
def test_scheduler_step_prediction_types(kind, pred):
    """sample / v_prediction heads and DDIM eta > 0 through the same fused kernel (only the coefficient row changes)"""
    import b200ldm
    ours = gc.make_scheduler(b200ldm, kind, 50, prediction_type=pred)
    ref = gc.make_scheduler(oracle, kind, 50, prediction_type=pred)
    x, m, nz = rnd((2, 4, 4, 4, 4), 54), rnd((2, 4, 4, 4, 4), 55) * 0.5, rnd((2, 4, 4, 4, 4), 56)
    kw = {} if kind == "ddpm" else {"eta": 0.5}
    for t in (int(ref.timesteps[0]), int(ref.timesteps[20]), int(ref.timesteps[-1])):
        p, p0 = ours.step(m.to(DEV), t, x.to(DEV), noise=nz.to(DEV), **kw)
        w, w0 = ref.step(m, t, x, noise=nz, **kw)
        assert relerr(p, w) <= 3e-5 and relerr(p0, w0) <= 3e-5, (t, relerr(p, w), relerr(p0, w0))


def test_scheduler_module_step_and_add_noise():
    import b200ldm
    ours, ref = gc.make_scheduler(b200ldm, "ddpm", 1000), gc.make_scheduler(oracle, "ddpm", 1000)
    x, e, nz = rnd((2, 4, 4, 4, 4), 51), rnd((2, 4, 4, 4, 4), 52), rnd((2, 4, 4, 4, 4), 53)
    for t in (999, 500, 0):
        p, p0 = ours.step(e.to(DEV), t, x.to(DEV), noise=nz.to(DEV))
        w, w0 = ref.step(e, t, x, noise=nz)
        assert relerr(p, w) <= 2e-5 and relerr(p0, w0) <= 2e-5
    tt = torch.tensor([3, 977])
    assert relerr(ours.add_noise(x.to(DEV), nz.to(DEV), tt.to(DEV)), ref.add_noise(x, nz, tt)) <= 1e-6


def test_philox_noise_is_standard_normal_and_matches_step_kernel():
    import b200ldm
    from b200ldm import ops
    z = ops.philox_normal(1 << 20, 1234, 7, DEV)
    assert abs(z.mean().item()) < 5e-3 and abs(z.std().item() - 1.0) < 5e-3
    assert abs((z ** 4).mean().item() - 3.0) < 0.1
    z2 = ops.philox_normal(1 << 20, 1234, 8, DEV)
    assert abs((z * z2).mean().item()) < 5e-3          # streams of different steps are uncorrelated
    # in-kernel noise == the standalone stream
    s = gc.make_scheduler(b200ldm, "ddpm", 10)
    coef = s.coef_table(DEV)
    n_el = 2 * 3 * 5 * 4 * 4
    counter = torch.full((1,), 3, dtype=torch.int32, device=DEV)
    x = ops.to_ndhwc(rnd((2, 4, 3, 5, 4), 61).to(DEV), 4, torch.float32)
    e = ops.to_ndhwc(rnd((2, 4, 3, 5, 4), 62).to(DEV), 4, torch.float32)
    xa, xb = x.clone(), x.clone()
    ops.scheduler_step(e, xa, coef, counter, 4, seed=99)
    nz = ops.philox_normal(n_el, 99, 3, DEV).reshape(x.shape)
    ops.scheduler_step(e, xb, coef, counter, 4, noise=nz)
    assert torch.equal(xa, xb)


def test_ae_sampling_and_mul_add():
    from b200ldm import ops
    mu, lv, nz = rnd((2, 3, 4, 5, 16), 71), rnd((2, 3, 4, 5, 16), 72, 20.0), rnd((2, 3, 4, 5, 16), 73)
    z, sig = ops.ae_sampling(mu.to(DEV), lv.to(DEV), nz.to(DEV))
    s_ref = torch.exp(lv.clamp(-30, 20) / 2)
    assert relerr(sig, s_ref) <= 1e-5 and relerr(z, mu + s_ref * nz) <= 1e-5
    assert relerr(ops.mul_add_f32(mu.to(DEV), lv.to(DEV), nz.to(DEV)), mu + lv * nz) <= 1e-6


def test_abi_is_reentrant_across_threads_and_streams():
    """SURVEY 8b "threading / streams": entry points are re-entrant across streams, the plan cache is mutex-guarded.
    Four host threads, each on its own CUDA stream, run conv (two kernels) + GroupNorm + attention chains concurrently;
    every thread must reproduce the single-threaded result."""
    import threading
    from b200ldm import ops
    dtype = torch.bfloat16
    x = to_act(q(rnd((2, 128, 9, 11, 7), 301), dtype), dtype)
    w1 = ops.repack_conv_weight((rnd((128, 128, 3, 3, 3), 302) * (128 * 27) ** -0.5).to(DEV), 128, 128, dtype)
    w2 = ops.repack_conv_weight((rnd((384, 128, 1, 1, 1), 303) * 128 ** -0.5).to(DEV), 384, 128, dtype)
    gamma, beta = (1.0 + 0.1 * rnd((128,), 304)).to(DEV), (0.1 * rnd((128,), 305)).to(DEV)

    def chain():
        h = ops.conv3d(x, w1, None, 128, 3, 1, 1, 1, want_stats=True)
        h = ops.groupnorm(h, gamma, beta, 32, 1e-6, True)
        qkv = ops.conv3d(h, w2, None, 384, ksize=1, stride=1, pad_lo=0)
        return ops.attention(qkv.reshape(2, 9 * 11 * 7, 384), 2, 64, 0.125).float()

    want = chain()          # also settles the measured conv kernel choices
    torch.cuda.synchronize()
    outs, errs = [None] * 4, []

    def worker(i):
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(6):
                    outs[i] = chain()
            s.synchronize()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for o in outs:
        assert relerr(o, want) <= 2e-2     # GroupNorm sums are atomics: not bitwise


@pytest.mark.parametrize("kernel", [2, 3], ids=["persistent", "w_halo"])
@pytest.mark.parametrize("case", [(2, (8, 16, 16), 64, 64, 2), (1, (9, 11, 13), 128, 128, 4), (2, (6, 8, 16), 128, 64, 2),
                                  (1, (8, 8, 8), 64, 128, 4)],
                         ids=lambda c: f"N{c[0]}_{'x'.join(map(str, c[1]))}_C{c[2]}->{c[3]}_w{c[4]}")
def test_conv3d_fine_chunk_stats_feed_groupnorm(case, kernel):
    """GroupNorm statistics per 2 / 4 channels (AutoencoderKL: 64 / 128 channels in 32 groups) from the persistent
    kernels' epilogues: the sums match the stored tensor and the following GroupNorm(+SiLU) needs no statistics pass."""
    import ctypes
    from b200ldm import _lib, ops
    lib = _lib.load()
    lib.b200ldm_debug_set_conv_kernel.argtypes = [ctypes.c_int]
    N, sp, C, Cout, cw = case
    x = q(rnd((N, C) + sp, 401), torch.bfloat16)
    w = q(rnd((Cout, C, 3, 3, 3), 402) * (C * 27) ** -0.5, torch.bfloat16)
    b = 0.1 * rnd((Cout,), 403)
    res = q(rnd((N, Cout) + sp, 404), torch.bfloat16)
    wp = ops.repack_conv_weight(w.to(DEV), Cout, C, torch.bfloat16)
    lib.b200ldm_debug_set_conv_kernel(kernel)
    try:
        y = ops.conv3d(to_act(x, torch.bfloat16), wp, b.to(DEV), Cout, 3, 1, 1, 1, residual=to_act(res, torch.bfloat16),
                       want_stats=True, stats_width=cw)
    finally:
        lib.b200ldm_debug_set_conv_kernel(-1)
    ref = F.conv3d(x.double(), w.double(), b.double(), padding=1) + res.double()
    assert relerr(from_act(y, Cout), ref) <= 1e-2
    cs = getattr(y, "cstats", None)
    assert cs is not None and y.cstats_width == cw and cs.numel() == N * (Cout // cw) * 2
    yf = from_act(y, Cout).double()
    want = torch.stack([yf.reshape(N, Cout // cw, cw, -1).sum(dim=(2, 3)), (yf ** 2).reshape(N, Cout // cw, cw, -1).sum(dim=(2, 3))], dim=-1)
    assert relerr(cs.reshape(N, Cout // cw, 2).cpu(), want) <= 1e-4
    groups = Cout // cw
    gamma, beta = 1.0 + 0.1 * rnd((Cout,), 405), 0.1 * rnd((Cout,), 406)
    launches = ops.launch_count()
    got = ops.groupnorm(y, gamma.to(DEV), beta.to(DEV), groups, 1e-6, True)
    assert ops.launch_count() - launches == 1                     # apply only: no statistics pass
    gref = F.silu(F.group_norm(yf, groups, gamma.double(), beta.double(), eps=1e-6))
    assert relerr(from_act(got, Cout), gref) <= 6e-3
