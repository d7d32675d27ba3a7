#!/usr/bin/env python
"""bench.py - headline benchmark of the latent-sampling hot path on B200 (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload unet_step_16g|sample_32g]

Default workload (N = 1 .. 8, weak scaling, one replica + one batch per GPU):
  BASELINE.json configs[1] - "BraTS-shaped" latent UNet denoise step: config_train_16g UNet
  (in 8 / out 8, channels (256,256,512), attention (F,T,T), 64-channel heads), batch 8, latent 36x44x28, bf16.
  One step = UNet noise prediction + fused DDPM scheduler update for the whole batch (one CUDA-graph replay).
  value = volume-steps/s (batch * steps / device time, max over ranks) with everything resident in HBM;
  e2e   = the same metric through the public module API (unet(x, t); scheduler.step) with pinned HOST buffers,
          the H2D copy of x_t and the D2H read of x_{t-1} inside the timed region.
  The same JSON line also carries "sampling": generated volumes/s of BASELINE.json configs[2] per GPU (full 1000-step
  DDPM sampling with the config_train_32g UNet at a 20^3 latent + AutoencoderKL decode to 80^3, batch 8 per GPU, final
  gather over NCCL) - the other half of BASELINE.json's metric (--no-sampling skips it, --sampling-T shortens it).
  Every throughput fraction is quoted twice: on ALGORITHMIC FLOPs and on EXECUTED FLOPs (`frac_executed`:
  the convs behind a nearest-x2 up-sampling run 8 of their 27 taps); the `sampling` object has its own `e2e` (host noise
  in, gathered volumes back in pinned host memory) and `roofline`.  "other_configs" carries BASELINE.json configs[3]
  (DDIM 50 steps at a 40^3 latent -> 160^3, 4 volumes per GPU) and configs[4] (AutoencoderKL encode + decode at
  96^3 .. 192^3), each resident and e2e (--no-extra skips them).
`--workload sample_32g`: that sampling workload alone as the headline line (--T steps, --latent edge, --batch volumes per
  call; `--scaling strong --total-volumes 64` shares BASELINE configs[2]'s 64 volumes between the GPUs).
`--impl reference`: the CPU oracle (pure-PyTorch restatement of the MONAI path; the reference itself is not
  installable here - DESIGN.md) timed on the host cores of the box with the GPU arm's config: as many of the batch's 8
  volumes as fit a 200 s budget for the requested steps (all 8 when they fit).
Default --steps: 100 denoise steps (a >= 2 s timed region, so the clock samples describe the region).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "3d-latent-diffusion-model_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

UNET_16G = dict(spatial_dims=3, in_channels=8, out_channels=8, channels=[256, 256, 512],
                attention_levels=[False, True, True], num_head_channels=[0, 64, 64], num_res_blocks=2)
UNET_32G = dict(spatial_dims=3, in_channels=32, out_channels=16, channels=[256, 512, 1024],
                attention_levels=[False, True, True], num_head_channels=[0, 64, 64], num_res_blocks=2)
AE_32G = dict(spatial_dims=3, in_channels=1, out_channels=1, latent_channels=16, channels=[64, 128, 256],
              num_res_blocks=2, norm_num_groups=32, norm_eps=1e-6, attention_levels=[False, False, True],
              with_encoder_nonlocal_attn=False, with_decoder_nonlocal_attn=False)
LATENT_16G = (36, 44, 28)   # diffusion patch [144,176,112] / 4 (3d_ldm/config/config_train_16g.json:52)
BATCH = 8
SCHED = dict(num_train_timesteps=1000, schedule="scaled_linear_beta", beta_start=0.0015, beta_end=0.0195)
FLOP_PER_VOLUME_STEP_16G = 2969.7e9   # SURVEY.md 8d / BASELINE.md 5 (algorithmic, unpadded)


# ---------------------------------------------------------------------------------------------- FLOP census
# SURVEY.md App. F counting rules: conv 2*vox*Cout*Cin*k^3, linear 2MKN, attention core 4*N^2*C; restated here so the
# bench can quote EXECUTED next to ALGORITHMIC work: the convs that follow a nearest-x2 up-sampling run 8 instead of 27
# taps per output voxel on the tensor-core path (phase decomposition), so 19/27 of their algorithmic
# FLOPs are never executed and must not be credited as throughput (VERDICT r01 bench hygiene #10).
def _conv_f(cin, cout, sp, taps=27):
    return 2.0 * math.prod(sp) * cin * cout * taps


def _attn_f(c, sp):
    n = math.prod(sp)
    return 8.0 * n * c * c + 4.0 * n * n * c


def _res_f(cin, cout, sp):
    return _conv_f(cin, cout, sp) + _conv_f(cout, cout, sp) + (_conv_f(cin, cout, sp, 1) if cin != cout else 0.0)


def unet_flops(kw, spatial, taps_up=27):
    """FLOPs of one DiffusionModelUNet forward for ONE volume (time-embedding GEMVs left out: < 0.1 %)"""
    ch, att = list(kw["channels"]), list(kw["attention_levels"])
    L = len(ch)
    nres = kw["num_res_blocks"]
    nres = list(nres) if isinstance(nres, (list, tuple)) else [nres] * L
    sps = [tuple(spatial)]
    for _ in range(L - 1):
        sps.append(tuple((v + 1) // 2 for v in sps[-1]))
    f = _conv_f(kw["in_channels"], ch[0], sps[0])
    skips, cout = [ch[0]], ch[0]
    for i in range(L):
        cin, cout = cout, ch[i]
        for j in range(nres[i]):
            f += _res_f(cin if j == 0 else cout, cout, sps[i]) + (_attn_f(cout, sps[i]) if att[i] else 0.0)
            skips.append(cout)
        if i != L - 1:
            f += _conv_f(cout, cout, sps[i + 1])
            skips.append(cout)
    f += 2 * _res_f(ch[-1], ch[-1], sps[-1]) + _attn_f(ch[-1], sps[-1])
    rc, rn, ra = ch[::-1], nres[::-1], att[::-1]
    cout = rc[0]
    for i in range(L):
        prev, cout = cout, rc[i]
        for j in range(rn[i] + 1):
            f += _res_f((prev if j == 0 else cout) + skips.pop(), cout, sps[L - 1 - i])
            f += _attn_f(cout, sps[L - 1 - i]) if ra[i] else 0.0
        if i != L - 1:
            f += _conv_f(cout, cout, sps[L - 2 - i], taps_up)
    return f + _conv_f(ch[0], kw["out_channels"], sps[0])


def decoder_flops(kw, latent_sp, taps_up=27):
    """FLOPs of AutoencoderKL.decode for ONE volume"""
    ch, att = list(kw["channels"]), list(kw["attention_levels"])
    L = len(ch)
    nres = kw["num_res_blocks"]
    nres = list(nres) if isinstance(nres, (list, tuple)) else [nres] * L
    rc, ra, rn = ch[::-1], att[::-1], nres[::-1]
    sp, lc = tuple(latent_sp), kw["latent_channels"]
    f = _conv_f(lc, lc, sp, 1) + _conv_f(lc, rc[0], sp)
    if kw.get("with_decoder_nonlocal_attn", True):
        f += 2 * _res_f(rc[0], rc[0], sp) + _attn_f(rc[0], sp)
    cout = rc[0]
    for i in range(L):
        cin, cout = cout, rc[i]
        for j in range(rn[i]):
            f += _res_f(cin if j == 0 else cout, cout, sp) + (_attn_f(cout, sp) if ra[i] else 0.0)
        if i != L - 1:
            sp = tuple(2 * v for v in sp)
            f += _conv_f(cout, cout, sp, taps_up)
    return f + _conv_f(cout, kw["out_channels"], sp)


def encoder_flops(kw, image_sp):
    """FLOPs of AutoencoderKL.encode for ONE volume (MONAI Encoder: conv_in, per level res blocks (+ attention) and a
    stride-2 conv between levels, optional non-local trio, conv to the latent channels, the two 1x1x1 quant convs)"""
    ch, att = list(kw["channels"]), list(kw["attention_levels"])
    L = len(ch)
    nres = kw["num_res_blocks"]
    nres = list(nres) if isinstance(nres, (list, tuple)) else [nres] * L
    sp, lc = tuple(image_sp), kw["latent_channels"]
    f = _conv_f(kw["in_channels"], ch[0], sp)
    cout = ch[0]
    for i in range(L):
        cin, cout = cout, ch[i]
        for j in range(nres[i]):
            f += _res_f(cin if j == 0 else cout, cout, sp) + (_attn_f(cout, sp) if att[i] else 0.0)
        if i != L - 1:
            sp = tuple((v + 1) // 2 for v in sp)
            f += _conv_f(cout, cout, sp)
    if kw.get("with_encoder_nonlocal_attn", True):
        f += 2 * _res_f(cout, cout, sp) + _attn_f(cout, sp)
    return f + _conv_f(cout, lc, sp) + 2 * _conv_f(lc, lc, sp, 1)


def executed_fraction_unet(kw, spatial):
    return unet_flops(kw, spatial, 8) / unet_flops(kw, spatial, 27)


def config_16g(world):
    """`config` of the headline line - the SAME dict on the GPU arm and on the --impl reference arm"""
    return {"workload": "unet_step_16g: config_train_16g UNet (256,256,512) denoise step + DDPM scheduler update, "
                        "BASELINE.json configs[1]",
            "batch_per_gpu": BATCH, "latent": list(LATENT_16G), "parallelism": f"replica-per-gpu x{world}",
            "l2": "inputs larger than L2 (weights 0.3 GB + 181 MB activations per layer vs 126 MB L2)"}


def seeded_init(model, seed):
    """Random-init weights of the named architecture (no checkpoints offline); every parameter is redrawn so the
    zero-initialised convs of the MONAI recipe do not make the network trivially zero."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.ndim >= 2:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) / math.sqrt(fan_in))
            elif name.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
    return model


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"], "hbm": p["hbm_gbs"],
                "src": "measured"}
    except Exception:
        return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "src": "fallback"}


def roofline_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the ncu --set full capture
    committed for THIS round (profiles/roofline_traffic.json, written by scripts/ncu_summary.py with the commit it was
    taken at); None when there is none - a number from an older build would not describe the kernels being benched."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        return float(d["traffic_bytes_per_launch"]), {k: d.get(k) for k in ("kernel", "launch", "algorithmic_bytes_per_launch",
                                                                            "source", "commit")}
    except Exception:
        return None, {"source": "no ncu --set full capture committed for this build (profiles/roofline_traffic.json)"}


def dist_setup():
    from b200ldm import dist as bd
    rank, world, local = bd.init_from_env()
    return rank, world, local


def timed_region(fn, steps, world):
    """barrier + sync, CUDA events on the launching stream around exactly `steps` calls, max over ranks."""
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


# ---------------------------------------------------------------------------------------------- CPU oracle legs
def cpu_oracle_unet_step(threads, budget_s, n_timed=1, n_warm=1):
    """The oracle's UNet-16g denoise step (+ scheduler step) on the host cores for a bounded sample of the workload:
    the largest number of the batch's 8 volumes (processed as ONE batch, full 36x44x28 latent) whose n_warm + n_timed
    steps fit `budget_s` (all 8 when they fit - then the sample IS the workload), else one volume cropped."""
    import oracle
    torch.set_num_threads(threads)
    unet = oracle.randomize_parameters(oracle.DiffusionModelUNet(**UNET_16G), 1234).eval()
    sched = oracle.DDPMScheduler(**SCHED)
    g = torch.Generator().manual_seed(0)

    def step(x):
        with torch.no_grad():
            eps = unet(x, torch.tensor([500.0]))
            sched.step(eps, 500, x)

    x1 = torch.randn((1, 8) + LATENT_16G, generator=g)
    step(x1)                                       # first touch (allocator, oneDNN primitive cache)
    t0 = time.perf_counter()
    step(x1)
    probe = time.perf_counter() - t0               # one volume, one step
    n_steps = n_timed + n_warm
    lat, frac, nvol = LATENT_16G, 1.0, 1
    for n in (8, 4, 2, 1):
        if probe * n * n_steps <= budget_s:
            nvol = n
            break
    else:
        lat, frac = (16, 24, 16), 16 * 24 * 16 / (36 * 44 * 28)
    x = torch.randn((nvol, 8) + lat, generator=g)
    times = []
    for i in range(n_steps):
        t0 = time.perf_counter()
        step(x)
        if i >= n_warm:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    if frac == 1.0:
        desc = (f"{nvol} of the {BATCH} volumes as one batch, full 36x44x28 latent" if nvol < BATCH
                else f"the whole batch of {BATCH} volumes, full 36x44x28 latent (the sample is the workload)")
    else:
        desc = "1 volume cropped to a 16x24x16 latent (13.9 % of the voxels; full-volume equivalents by voxel count)"
    return {"value": frac * nvol * len(times) / total, "unit": "volume-steps/s", "cores": threads, "kind": "port",
            "sample": desc + f", UNet forward + scheduler step; {len(times)} timed steps after {n_warm} warm-up, "
                             f"{total:.1f} s of CPU work"}, total / len(times)


def run_reference(args):
    """--impl reference: the reference's own (CPU) implementation of the path.  MONAI is not installable offline, so this
    is the oracle port (cpu_baseline.kind "port"), all host threads, on the GPU arm's config / metric / unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    budget = float(os.environ.get("B200LDM_BENCH_REF_BUDGET_S", "200"))
    cb, s_step = cpu_oracle_unet_step(threads, budget_s=budget, n_timed=args.steps, n_warm=max(1, args.warmup))
    line = {"impl": "reference", "metric": "unet_denoise_volume_steps_per_sec", "value": cb["value"],
            "unit": "volume-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": max(1, args.warmup),
            "ms_per_step": s_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_16g(args.gpus), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "volume-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def cpu_oracle_cfg1(threads):
    """BASELINE.json configs[0] exactly, on the host cores: config_train_32g UNet + AutoencoderKL, ONE 16^3 latent (64^3
    volume), 10 DDPM steps + decode (the inference.py flow, concat conditioning), fp32 oracle, single shot after one
    warm-up UNet step.  5.54 TFLOP of algorithmic work."""
    import oracle
    torch.set_num_threads(threads)
    unet = oracle.randomize_parameters(oracle.DiffusionModelUNet(**UNET_32G), 1234).eval()
    ae = oracle.randomize_parameters(oracle.AutoencoderKL(**AE_32G), 4321).eval()
    sched = oracle.DDPMScheduler(**SCHED)
    sched.set_timesteps(10)
    g = torch.Generator().manual_seed(42)
    x = torch.randn((1, 16, 16, 16, 16), generator=g)
    c = torch.randn((1, 16, 16, 16, 16), generator=g)
    with torch.no_grad():
        unet(torch.cat([x, c], 1), torch.tensor([900.0]))           # warm-up step (allocator, oneDNN primitives)
        t0 = time.perf_counter()
        oracle.LatentDiffusionInferer(sched, 1.0).sample(x, ae, unet, sched, conditioning=c, mode="concat", verbose=False)
        dt = time.perf_counter() - t0
    out = {"value": 1.0 / dt, "unit": "volumes/s", "seconds_per_volume": dt, "cores": threads, "kind": "port",
           "sample": "the whole of configs[0]: 1 volume, 10 DDPM steps at a 16^3 latent + decode to 64^3, single shot"}
    if threads > 4:
        # the reference's own setting (3d_ldm/inference.py:58: torch.set_num_threads(4)); SURVEY.md 8d asks for both
        torch.set_num_threads(4)
        with torch.no_grad():
            t0 = time.perf_counter()
            oracle.LatentDiffusionInferer(sched, 1.0).sample(x, ae, unet, sched, conditioning=c, mode="concat", verbose=False)
            out["seconds_per_volume_4_threads"] = time.perf_counter() - t0
        torch.set_num_threads(threads)
    return out


# ---------------------------------------------------------------------------------------------- GPU workloads
def profile_eager_steps(body, n_prof=8, n_heat=8):
    """per-kernel CUDA events (on the launching stream) around every launch of `n_prof` eager steps; returns
    (kinds, by_tag, ms per step summed over kernels).  `body` must already have run once (weights packed, convs tuned).
    `n_heat` unprofiled steps run first so the profiled ones see the clocks of a chip under sustained load (sw_power_cap),
    like the timed region - two cold steps used to read ~5 % faster than the replayed graph for that reason alone."""
    from b200ldm import ops
    for _ in range(n_heat):
        body()
    torch.cuda.synchronize()
    ops.PROFILER = ops.KernelProfiler()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_prof):
        body()
    e1.record()
    torch.cuda.synchronize()
    kinds, by_tag = ops.PROFILER.summary()
    ops.PROFILER = None
    profile_eager_steps.last_wall_ms = e0.elapsed_time(e1) / n_prof    # eager step incl. launch gaps and event records
    for d in (kinds, by_tag):
        for v in d.values():
            v["ms"] /= n_prof
            v["work"] /= n_prof
            v["launches"] //= n_prof
    return kinds, by_tag, sum(v["ms"] for v in kinds.values())


def conv_roofline(kinds, by_tag, step_ms, pk, kernel_desc):
    """roofline object of the dominant kernel family (the tcgen05 implicit-GEMM convs of one step): `achieved` =
    EXECUTED FLOPs of all conv launches of a step / their summed CUDA-event durations"""
    conv = kinds["conv3d"]
    conv_tf = conv["work"] / (conv["ms"] * 1e-3) / 1e12
    traffic, tsrc = roofline_traffic()
    roof = {"bound": "tensor", "kernel": kernel_desc, "achieved": conv_tf, "peak": pk["bf16_sustained"],
            "unit": "TFLOP/s", "frac": conv_tf / pk["bf16_sustained"],
            "flops": "executed (fused nearest-x2 up-sampling convs counted with the 8 taps they run, not 27)",
            "peak_source": f"{pk['src']} MEASURED_PEAKS.json bf16_tflops_sustained (kernels timed inside a full step); "
                           f"burst {pk['bf16_burst']:.0f}", "frac_of_burst": conv_tf / pk["bf16_burst"],
            "traffic": traffic, "traffic_source": tsrc,
            "launches_per_step": conv["launches"], "ms_per_step": conv["ms"], "share_of_step": conv["ms"] / step_ms}
    top = sorted(((v["ms"], k) for k, v in by_tag.items() if k[0] == "conv3d"), reverse=True)[:4]
    roof["top_shapes"] = [{"shape": k[1], "ms": by_tag[k]["ms"], "launches": by_tag[k]["launches"],
                           "TFLOPs": by_tag[k]["work"] / (by_tag[k]["ms"] * 1e-3) / 1e12} for _, k in top]
    return roof


def kernel_breakdown(kinds, pk, graph_ms=None):
    out = {k: {"ms_per_step": v["ms"], "launches_per_step": v["launches"]} for k, v in kinds.items()}
    for k in ("gn_stats", "gn_apply", "gn_small", "tap_gather"):
        if k in kinds and kinds[k]["ms"] > 0:
            out[k]["GBps"] = kinds[k]["work"] / (kinds[k]["ms"] * 1e-3) / 1e9
            out[k]["frac_hbm"] = out[k]["GBps"] / pk["hbm"]
    for k in ("attention", "conv3d"):
        if k in kinds:
            out[k]["TFLOPs"] = kinds[k]["work"] / (kinds[k]["ms"] * 1e-3) / 1e12
    total = sum(v["ms"] for v in kinds.values())
    out["_sum_of_kernels_ms"] = total
    out["_eager_step_wall_ms"] = getattr(profile_eager_steps, "last_wall_ms", None)
    if graph_ms is not None:
        # what the per-kernel events do not see: helper kernels without a profiler hook (select_row, counter_add, fills),
        # launch gaps, and - with sub-batch streams - the overlap the graph gains (negative = the graph is faster)
        out["_graph_replay_ms"] = graph_ms
        out["_unattributed_ms"] = graph_ms - total
    return out


def bench_unet_step(args, rank, world, local):
    import b200ldm
    from b200ldm.inferer import SamplingLoop
    dev = torch.device("cuda", local)
    unet = seeded_init(b200ldm.DiffusionModelUNet(**UNET_16G, precision="bf16"), 1234).to(dev).eval()
    sched = b200ldm.DDPMScheduler(**SCHED)
    g = torch.Generator().manual_seed(42 + 1000 * rank)
    x_T = torch.randn((BATCH, 8) + LATENT_16G, generator=g)
    warm = max(args.warmup, 3)
    if args.steps + warm + 2 > 1000:
        raise SystemExit("steps + warmup must stay below the 1000-step DDPM schedule")

    # ---- device-resident leg: graph-replayed denoise steps
    loop = SamplingLoop(unet, sched, x_T.to(dev), None, "crossattn", seed=42 + 1000 * rank, use_graph=True)
    loop.step()                                   # capture (eager warm-up inside) - counts as warm-up step 1
    per_graph = loop.launches_per_step            # kernels recorded into the graph = launched per replay
    loop_subs = list(loop.subs)                   # sub-batches of the top UNet level (parallel graph branches)
    for _ in range(warm - 1):
        loop.step()
    with ClockSampler(local) as clk:
        ms = timed_region(loop.step, args.steps, world)
    value = world * BATCH * args.steps / (ms * 1e-3)

    # ---- end-to-end leg: public module API, pinned host buffers, copies inside the timed region
    x_host = x_T.clone().pin_memory()
    out_host = torch.empty_like(x_host).pin_memory()
    t_dev = torch.tensor([500.0], device=dev)

    def e2e_step():
        x = x_host.to(dev, non_blocking=True)
        eps = unet(x, t_dev)
        prev, _ = sched.step(eps, 500, x)
        out_host.copy_(prev, non_blocking=True)

    for _ in range(3):
        e2e_step()
    e2e_steps = max(3, min(args.steps, 50))
    ms_e2e = timed_region(e2e_step, e2e_steps, world)
    e2e_value = world * BATCH * e2e_steps / (ms_e2e * 1e-3)
    nbytes = x_host.numel() * 4

    # ---- roofline leg: per-kernel CUDA events inside eager steps (a single-stream loop: per-kernel events mean nothing
    # while two sub-batch streams share the SMs)
    pk = peaks()
    prof_loop = SamplingLoop(unet, sched, x_T.to(dev), None, "crossattn", seed=42 + 1000 * rank, use_graph=False)
    prof_loop._body()
    kinds, by_tag, step_ms_prof = profile_eager_steps(prof_loop._body)
    del prof_loop
    roof = conv_roofline(kinds, by_tag, step_ms_prof, pk,
                         "tcgen05 implicit-GEMM conv kernels (conv3d_tc / conv3d_tcp / conv3d_halo), all conv launches "
                         "of one denoise step")
    breakdown = kernel_breakdown(kinds, pk, ms / args.steps)

    # ---- second half of BASELINE.json's metric: generated volumes/s (full sampling + decode)
    sampling = None
    if not args.no_sampling:
        del loop
        torch.cuda.empty_cache()
        sampling = run_sampling(rank, world, local, args.sampling_T, 20, reps=1, warm=1, batch=BATCH,
                                total=BATCH * world)

    extra = None
    if not args.no_extra:
        torch.cuda.empty_cache()
        extra = run_big_configs(rank, world, local)

    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_oracle_unet_step(os.cpu_count() or 1, budget_s=30.0, n_timed=1, n_warm=1)
        if extra is not None:
            # the reference's own CPU-runnable case, timed on the same host cores next to the GPU's number for it
            extra["cfg1_default_inference"]["cpu_baseline"] = cpu_oracle_cfg1(os.cpu_count() or 1)
    step_s = ms / args.steps * 1e-3
    step_tf = FLOP_PER_VOLUME_STEP_16G * BATCH / step_s / 1e12
    fx = executed_fraction_unet(UNET_16G, LATENT_16G)
    line = {"metric": "unet_denoise_volume_steps_per_sec", "value": value, "unit": "volume-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config_16g(world), "impl_detail": {"cuda_graph": True, "batch_streams": len(loop_subs)},
            "step_tflops_algorithmic": step_tf, "frac_of_sustained_peak": step_tf / pk["bf16_sustained"],
            "step_tflops_executed": step_tf * fx, "frac_executed": step_tf * fx / pk["bf16_sustained"],
            "executed_over_algorithmic_flops": fx,
            "e2e": {"value": e2e_value, "unit": "volume-steps/s", "h2d_bytes_per_step": nbytes,
                    "d2h_bytes_per_step": nbytes, "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps},
            "gpu_launches": per_graph * args.steps, "launches_per_step": per_graph,
            "roofline": roof, "kernel_breakdown": breakdown, "sampling": sampling, "other_configs": extra,
            "cpu_baseline": cpu,
            "clocks": clk.summary()}
    print(json.dumps(line))


FLOP_UNET_32G = {16: 465.6e9, 20: 914.5e9, 40: 7910.7e9}     # per volume per step
FLOP_DEC_32G = {16: 888.7e9, 20: 1799.7e9, 40: 21738e9}      # per volume


def run_sampling(rank, world, local, T, latent, reps, warm, batch=BATCH, total=None, want_e2e=True, want_roofline=True):
    """BASELINE.json configs[2]: full DDPM latent sampling (config_train_32g UNet, concat conditioning) + AutoencoderKL
    decode + final NCCL gather of volumes and seeds.  `total` volumes are sharded over the ranks (weak scaling:
    total = batch * world; strong scaling: total fixed) and every rank works through its shard in batches of `batch`.
    Returns the `sampling` object of the bench line (rank 0; timing = max over ranks)."""
    import b200ldm
    from b200ldm import dist as bd, ops
    from b200ldm.inferer import SamplingLoop
    dev = torch.device("cuda", local)
    unet = seeded_init(b200ldm.DiffusionModelUNet(**UNET_32G, precision="bf16"), 1234).to(dev).eval()
    ae = seeded_init(b200ldm.AutoencoderKL(**AE_32G, precision="bf16"), 4321).to(dev).eval()
    sched = b200ldm.DDPMScheduler(**SCHED)
    sched.set_timesteps(T)
    inf = b200ldm.LatentDiffusionInferer(sched, scale_factor=1.0)
    lat = (16, latent, latent, latent)
    total = total or batch * world
    g = torch.Generator().manual_seed(7)
    cond_host = torch.randn((total, 16) + lat[1:], generator=g).pin_memory()
    noise_host = torch.randn((total,) + lat, generator=g).pin_memory()
    cond_dev, noise_dev = cond_host.to(dev), noise_host.to(dev)
    out_host = torch.empty((total, 1, 4 * latent, 4 * latent, 4 * latent), dtype=torch.float32).pin_memory()
    seed = [42]

    def resident():   # inputs already in HBM, result left in HBM
        seed[0] += 1   # a new in-kernel noise stream every call
        return bd.sample_sharded(inf, ae, unet, sched, total, lat, base_seed=seed[0], conditioning=cond_dev,
                                 mode="concat", device=dev, noise=noise_dev, max_batch=batch)

    def e2e():        # what inference.py does around the path: host noise in, volumes back on the host
        seed[0] += 1
        vols, _ = bd.sample_sharded(inf, ae, unet, sched, total, lat, base_seed=seed[0], conditioning=cond_host,
                                    mode="concat", device=dev, noise=noise_host, max_batch=batch)
        if rank == 0:
            out_host.copy_(vols, non_blocking=True)

    for _ in range(warm):
        resident()
    l0 = ops.launch_count()
    with ClockSampler(local) as clk:
        ms = timed_region(resident, reps, world) / reps
    launches = (ops.launch_count() - l0) // reps
    pk = peaks()
    vps = total / (ms * 1e-3)
    out = {"metric": "generated_volumes_per_sec", "value": vps, "unit": "volumes/s", "ms_per_batch": ms,
           "workload": f"sample_32g: DDPM {T}-step latent sampling + AutoencoderKL decode, latent {latent}^3 -> "
                       f"{4 * latent}^3, concat conditioning, {total} volumes over {world} GPU(s) in batches of {batch}, "
                       "final gather of volumes + seeds (BASELINE.json configs[2])",
           "volumes_total": total, "batch": batch, "gpu_launches": launches, "clocks": clk.summary()}
    if latent in FLOP_UNET_32G:
        per_gpu = total / world
        flop = per_gpu * (T * FLOP_UNET_32G[latent] + FLOP_DEC_32G[latent])
        fx_u = executed_fraction_unet(UNET_32G, (latent,) * 3)
        fx_d = decoder_flops(AE_32G, (latent,) * 3, 8) / decoder_flops(AE_32G, (latent,) * 3, 27)
        flop_x = per_gpu * (T * FLOP_UNET_32G[latent] * fx_u + FLOP_DEC_32G[latent] * fx_d)
        out["tflops_algorithmic"] = flop / (ms * 1e-3) / 1e12
        out["frac_of_sustained_peak"] = out["tflops_algorithmic"] / pk["bf16_sustained"]
        out["tflops_executed"] = flop_x / (ms * 1e-3) / 1e12
        out["frac_executed"] = out["tflops_executed"] / pk["bf16_sustained"]
    if want_e2e:
        e2e()
        ms_e2e = timed_region(e2e, 1, world)
        out["e2e"] = {"value": total / (ms_e2e * 1e-3), "unit": "volumes/s", "ms_per_batch": ms_e2e,
                      "h2d_bytes_per_step": (noise_host.numel() + cond_host.numel()) * 4,
                      "d2h_bytes_per_step": out_host.numel() * 4,
                      "what": "x_T and conditioning from pinned host memory, gathered volumes copied into pinned host "
                              "memory on rank 0, all inside the timed region"}
    if want_roofline:
        # dominant kernel of the sampling workload: the convs of the 32g denoise step (T x per-step vs one decode)
        nb = min(batch, (total + world - 1) // world)
        pl = SamplingLoop(unet, sched, noise_dev[:nb], cond_dev[:nb], "concat", seed=1, use_graph=False)
        pl._body()
        kinds, by_tag, step_ms = profile_eager_steps(pl._body)
        out["roofline"] = conv_roofline(kinds, by_tag, step_ms, pk,
                                        "tcgen05 implicit-GEMM conv kernels, all conv launches of one 32g denoise step")
        out["kernel_breakdown"] = kernel_breakdown(kinds, pk)
    return out


def run_big_configs(rank, world, local, ddim_batch=4, edges=(96, 128, 160, 192)):
    """BASELINE.json configs[3] and configs[4] per GPU, through the public API (weak scaling: every rank does the same):
    * `ddim50_40`: DDIM 50-step sampling with the config_train_32g UNet at a 40^3 latent + AutoencoderKL decode to 160^3,
      `ddim_batch` volumes per GPU (configs[3]: 32 volumes over 8 GPUs) - the attention-heavy path (L = 8000 / 1000 tokens
      in the UNet, L = 64 000 with 256-channel heads in the decoder);
    * `ae_sweep`: AutoencoderKL `reconstruct` (encode -> mu -> decode) of one volume per GPU at each edge.
    Each with a device-resident number and an e2e number (pinned host input, result back in pinned host memory)."""
    import b200ldm
    dev = torch.device("cuda", local)
    pk = peaks()
    unet = seeded_init(b200ldm.DiffusionModelUNet(**UNET_32G, precision="bf16"), 1234).to(dev).eval()
    ae = seeded_init(b200ldm.AutoencoderKL(**AE_32G, precision="bf16"), 4321).to(dev).eval()
    out = {}
    # ---- configs[0] on the GPU: the reference's own CPU-runnable case (inference.py defaults), latency of one volume
    s10 = b200ldm.DDPMScheduler(**SCHED)
    s10.set_timesteps(10)
    inf10 = b200ldm.LatentDiffusionInferer(s10, 1.0)
    g1 = torch.Generator().manual_seed(42)
    x1_host = torch.randn((1, 16, 16, 16, 16), generator=g1).pin_memory()
    c1_host = torch.randn((1, 16, 16, 16, 16), generator=g1).pin_memory()
    v1_host = torch.empty((1, 1, 64, 64, 64)).pin_memory()

    def cfg1_e2e():
        v1_host.copy_(inf10.sample(input_noise=x1_host.to(dev, non_blocking=True), autoencoder_model=ae, diffusion_model=unet,
                                   scheduler=s10, conditioning=c1_host.to(dev, non_blocking=True), mode="concat",
                                   verbose=False), non_blocking=True)

    cfg1_e2e()
    cfg1_e2e()
    ms1 = timed_region(cfg1_e2e, 10, world) / 10
    out["cfg1_default_inference"] = {
        "metric": "generated_volumes_per_sec", "value": world * 1e3 / ms1, "unit": "volumes/s", "ms_per_volume": ms1,
        "workload": "config_train_32g UNet + AutoencoderKL, ONE 16^3 latent -> 64^3 volume, 10 DDPM steps + decode through "
                    "LatentDiffusionInferer.sample, host noise in / host volume out inside the timed region "
                    "(BASELINE.json configs[0]; 5.54 TFLOP algorithmic)",
        "tflops_algorithmic": (10 * FLOP_UNET_32G[16] + FLOP_DEC_32G[16]) / (ms1 * 1e-3) / 1e12,
        "h2d_bytes_per_step": (x1_host.numel() + c1_host.numel()) * 4, "d2h_bytes_per_step": v1_host.numel() * 4,
        "parity": "tests/test_gpu_models.py::test_cfg1_default_config_sampling_psnr (>= 40 dB vs the CPU oracle)"}
    # ---- configs[3]
    sched = b200ldm.DDIMScheduler(**SCHED)
    sched.set_timesteps(50)
    inf = b200ldm.LatentDiffusionInferer(sched, 1.0)
    g = torch.Generator().manual_seed(11 + rank)
    lat = (16, 40, 40, 40)
    x_host = torch.randn((ddim_batch,) + lat, generator=g).pin_memory()
    c_host = torch.randn((ddim_batch,) + lat, generator=g).pin_memory()
    v_host = torch.empty((ddim_batch, 1, 160, 160, 160)).pin_memory()
    x_dev, c_dev = x_host.to(dev), c_host.to(dev)

    def sample(x, c):
        return inf.sample(input_noise=x, autoencoder_model=ae, diffusion_model=unet, scheduler=sched, conditioning=c,
                          mode="concat", verbose=False)

    def ddim_e2e():
        v_host.copy_(sample(x_host.to(dev, non_blocking=True), c_host.to(dev, non_blocking=True)), non_blocking=True)

    sample(x_dev, c_dev)                       # convs measured, loop + decode graphs captured
    with ClockSampler(local) as clk:
        ms = timed_region(lambda: sample(x_dev, c_dev), 2, world) / 2
    ddim_e2e()
    ms_e2e = timed_region(ddim_e2e, 1, world)
    flop = ddim_batch * (50 * FLOP_UNET_32G[40] + FLOP_DEC_32G[40])
    fx = ddim_batch * (50 * FLOP_UNET_32G[40] * executed_fraction_unet(UNET_32G, (40,) * 3)
                       + decoder_flops(AE_32G, (40,) * 3, 8))
    out["ddim50_40"] = {
        "metric": "generated_volumes_per_sec", "value": world * ddim_batch / (ms * 1e-3), "unit": "volumes/s",
        "ms_per_batch": ms, "batch_per_gpu": ddim_batch,
        "workload": f"DDIM 50-step sampling (config_train_32g UNet, concat conditioning) at a 40^3 latent + AutoencoderKL "
                    f"decode to 160^3, {ddim_batch} volumes per GPU x {world} GPU(s) (BASELINE.json configs[3])",
        "tflops_algorithmic": flop / (ms * 1e-3) / 1e12, "frac_of_sustained_peak": flop / (ms * 1e-3) / 1e12 / pk["bf16_sustained"],
        "tflops_executed": fx / (ms * 1e-3) / 1e12, "frac_executed": fx / (ms * 1e-3) / 1e12 / pk["bf16_sustained"],
        "e2e": {"value": world * ddim_batch / (ms_e2e * 1e-3), "unit": "volumes/s", "ms_per_batch": ms_e2e,
                "h2d_bytes_per_step": (x_host.numel() + c_host.numel()) * 4, "d2h_bytes_per_step": v_host.numel() * 4},
        "parity": "tests/test_gpu_long.py::test_cfg4_ddim_50_steps_40cubed_vs_oracle (40.3 dB, profiles/r02_long_parity)",
        "clocks": clk.summary()}
    del x_dev, c_dev, v_host
    inf.clear_cache() if hasattr(inf, "clear_cache") else None
    torch.cuda.empty_cache()
    # ---- configs[4]
    sweep = []
    for edge in edges:
        vol_host = torch.rand((1, 1, edge, edge, edge), generator=g).pin_memory()
        rec_host = torch.empty_like(vol_host).pin_memory()
        vol = vol_host.to(dev)

        def e2e_rec():
            rec_host.copy_(ae.reconstruct(vol_host.to(dev, non_blocking=True)), non_blocking=True)

        ae.reconstruct(vol)
        ae.reconstruct(vol)
        reps = 5 if edge <= 128 else 3
        ms = timed_region(lambda: ae.reconstruct(vol), reps, world) / reps
        e2e_rec()
        ms_e2e = timed_region(e2e_rec, reps, world) / reps
        fa = encoder_flops(AE_32G, (edge,) * 3) + decoder_flops(AE_32G, (edge // 4,) * 3, 27)
        fe = encoder_flops(AE_32G, (edge,) * 3) + decoder_flops(AE_32G, (edge // 4,) * 3, 8)
        sweep.append({"edge": edge, "ms": ms, "volumes_per_s": world / (ms * 1e-3),
                      "tflops_algorithmic": fa / (ms * 1e-3) / 1e12, "tflops_executed": fe / (ms * 1e-3) / 1e12,
                      "frac_executed": fe / (ms * 1e-3) / 1e12 / pk["bf16_sustained"],
                      "e2e_ms": ms_e2e, "e2e_volumes_per_s": world / (ms_e2e * 1e-3),
                      "h2d_bytes_per_step": vol_host.numel() * 4, "d2h_bytes_per_step": rec_host.numel() * 4})
        del vol, vol_host, rec_host
        torch.cuda.empty_cache()
    out["ae_sweep"] = {"metric": "autoencoder_reconstruct_volumes_per_sec", "unit": "volumes/s",
                       "workload": f"AutoencoderKL (config_train_32g) encode + decode of one volume per GPU x {world} GPU(s), "
                                   "bf16 (BASELINE.json configs[4])",
                       "parity": "tests/test_gpu_long.py::test_cfg5_autoencoder_reconstruct_vs_oracle (96^3 / 128^3: 53.7 / 54.5 dB)",
                       "sizes": sweep}
    return out


def bench_sample(args, rank, world, local):
    total = args.total_volumes if args.scaling == "strong" else args.batch * world
    sm = run_sampling(rank, world, local, args.T, args.latent, args.steps, max(1, args.warmup), batch=args.batch,
                      total=total)
    if rank != 0:
        return
    line = {"metric": sm["metric"], "value": sm["value"], "unit": sm["unit"], "n_gpus": world, "steps": args.steps,
            "warmup": max(1, args.warmup), "ms_per_step": sm["ms_per_batch"], "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": sm["workload"], "volumes_total": total, "batch": args.batch,
                       "parallelism": f"batch-sharded x{world}",
                       "l2": "inputs larger than L2 (1.3 GB of packed weights per step vs 126 MB L2)"},
            "tflops_algorithmic": sm.get("tflops_algorithmic"), "frac_of_sustained_peak": sm.get("frac_of_sustained_peak"),
            "tflops_executed": sm.get("tflops_executed"), "frac_executed": sm.get("frac_executed"),
            "e2e": sm.get("e2e"), "roofline": sm.get("roofline"), "kernel_breakdown": sm.get("kernel_breakdown"),
            "gpu_launches": sm["gpu_launches"] * args.steps, "clocks": sm["clocks"]}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None,
                    help="timed steps (default: 100 denoise steps = a >= 2 s timed region; sample_32g: 1 batch)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="unet_step_16g", choices=["unet_step_16g", "sample_32g"])
    ap.add_argument("--T", type=int, default=50, help="sample_32g: number of DDPM steps")
    ap.add_argument("--latent", type=int, default=20, help="sample_32g: latent edge")
    ap.add_argument("--batch", type=int, default=BATCH, help="sample_32g: volumes per sampling call on one GPU")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="sample_32g: weak = --batch volumes per GPU; strong = --total-volumes shared by all GPUs")
    ap.add_argument("--total-volumes", type=int, default=64, help="sample_32g --scaling strong (BASELINE configs[2]: 64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sampling", action="store_true", help="skip the volumes/s leg of the default workload")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the BASELINE configs[3] (DDIM-50 @ 40^3) and configs[4] (AutoencoderKL sweep) legs")
    ap.add_argument("--sampling-T", type=int, default=1000, help="DDPM steps of the volumes/s leg (reference: 1000)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 1 if args.workload == "sample_32g" else (100 if args.impl == "b200" else 3)
    if args.impl == "reference":
        return run_reference(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (b200ldm has no CPU fallback); use --impl reference for the CPU oracle")
    rank, world, local = dist_setup()
    from b200ldm import _lib
    _lib.check(_lib.load().b200ldm_check_device(), "check_device")
    if args.workload == "unet_step_16g":
        bench_unet_step(args, rank, world, local)
    else:
        bench_sample(args, rank, world, local)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
