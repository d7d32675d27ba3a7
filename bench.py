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
`--workload sample_32g`: that sampling workload alone as the headline line (--T steps, --latent edge).
`--impl reference`: the CPU oracle (pure-PyTorch restatement of the MONAI path; the reference itself is not
  installable here - DESIGN.md) timed on the host cores of the box on a bounded sample of the same workload.
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
    """Time the oracle UNet-16g denoise step (+ scheduler step) on the host for a bounded sample."""
    import oracle
    torch.set_num_threads(threads)
    unet = oracle.randomize_parameters(oracle.DiffusionModelUNet(**UNET_16G), 1234).eval()
    sched = oracle.DDPMScheduler(**SCHED)
    candidates = [(LATENT_16G, 1.0, "1 of the 8 volumes, full 36x44x28 latent, 1 denoise step"),
                  ((16, 24, 16), 16 * 24 * 16 / (36 * 44 * 28), "1 volume cropped to a 16x24x16 latent (13.9% of the voxels), "
                                                                "1 denoise step")]
    g = torch.Generator().manual_seed(0)
    for lat, frac, desc in candidates:
        x = torch.randn((1, 8) + lat, generator=g)
        t0 = time.perf_counter()
        with torch.no_grad():
            eps = unet(x, torch.tensor([500.0]))
            sched.step(eps, 500, x)
        probe = time.perf_counter() - t0
        if probe * (n_timed + n_warm) <= budget_s or lat == candidates[-1][0]:
            break
    times = []
    for i in range(n_warm - 1 + n_timed):
        t0 = time.perf_counter()
        with torch.no_grad():
            eps = unet(x, torch.tensor([500.0]))
            sched.step(eps, 500, x)
        dt = time.perf_counter() - t0
        if i >= n_warm - 1:
            times.append(dt)
    total = sum(times)
    return {"value": frac * len(times) / total, "unit": "volume-steps/s", "cores": threads, "kind": "port",
            "sample": desc + f" (full-volume equivalents by voxel count); {len(times)} timed"}, total / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    cb, ms_step = cpu_oracle_unet_step(threads, budget_s=200.0, n_timed=args.steps, n_warm=max(1, args.warmup))
    line = {"impl": "reference", "metric": "unet_denoise_volume_steps_per_sec", "value": cb["value"],
            "unit": "volume-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "unet_step_16g: config_train_16g UNet denoise step, latent 36x44x28 (CPU oracle, "
                                   "bounded sample)", "batch_per_gpu": BATCH},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "volume-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- GPU workloads
def bench_unet_step(args, rank, world, local):
    import b200ldm
    from b200ldm import ops
    from b200ldm.inferer import SamplingLoop
    dev = torch.device("cuda", local)
    unet = seeded_init(b200ldm.DiffusionModelUNet(**UNET_16G, precision="bf16"), 1234).to(dev).eval()
    sched = b200ldm.DDPMScheduler(**SCHED)
    g = torch.Generator().manual_seed(42 + 1000 * rank)
    x_T = torch.randn((BATCH, 8) + LATENT_16G, generator=g)
    if args.steps + args.warmup + 2 > 1000:
        raise SystemExit("steps + warmup must stay below the 1000-step DDPM schedule")

    # ---- device-resident leg: graph-replayed denoise steps
    loop = SamplingLoop(unet, sched, x_T.to(dev), None, "crossattn", seed=42 + 1000 * rank, use_graph=True)
    loop.step()                                   # capture (eager warm-up inside) - counts as warm-up step 1
    per_graph = loop.launches_per_step            # kernels recorded into the graph = launched per replay
    loop_subs = list(loop.subs)                   # sub-batches of the top UNet level (parallel graph branches)
    for _ in range(max(args.warmup, 3) - 1):
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

    for _ in range(2):
        e2e_step()
    e2e_steps = max(3, min(args.steps, 10))
    ms_e2e = timed_region(e2e_step, e2e_steps, world)
    e2e_value = world * BATCH * e2e_steps / (ms_e2e * 1e-3)
    nbytes = x_host.numel() * 4

    # ---- roofline leg: per-kernel CUDA events inside an eager step (same stream the kernels launch on)
    pk = peaks()
    # (a single-stream loop: per-kernel events mean nothing while two sub-batch streams share the SMs)
    prof_loop = SamplingLoop(unet, sched, x_T.to(dev), None, "crossattn", seed=42 + 1000 * rank, use_graph=False)
    prof_loop._body()
    torch.cuda.synchronize()
    ops.PROFILER = ops.KernelProfiler()
    n_prof = 2
    for _ in range(n_prof):
        prof_loop._body()
    torch.cuda.synchronize()
    del prof_loop
    kinds, by_tag = ops.PROFILER.summary()
    ops.PROFILER = None
    conv = kinds["conv3d"]
    conv_tf = conv["work"] / (conv["ms"] * 1e-3) / 1e12
    step_ms_prof = sum(v["ms"] for v in kinds.values()) / n_prof
    roof = {"bound": "tensor", "kernel": "conv3d_tc_kernel (tcgen05 implicit GEMM, all conv launches of one step)",
            "achieved": conv_tf, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": conv_tf / pk["bf16_sustained"],
            "peak_source": f"{pk['src']} (sustained: kernels timed inside a full step)",
            # dram__bytes_read.sum + dram__bytes_write.sum of the dominant launch (256->256 k3 @8x36x44x28, 844 us,
            # tensor pipe 83.5 %) from the committed ncu --set full capture; algorithmic in + out = 363 MB
            "traffic": 330.2e6, "traffic_source": "profiles/r01b_ncu_full.md (per launch of conv3d_tc_kernel<256,2,1>)",
            "launches_per_step": conv["launches"] // n_prof, "share_of_step": conv["ms"] / n_prof / step_ms_prof}
    breakdown = {k: {"ms_per_step": v["ms"] / n_prof, "launches_per_step": v["launches"] // n_prof} for k, v in kinds.items()}
    for k in ("gn_stats", "gn_apply"):
        if k in kinds:
            breakdown[k]["GBps"] = kinds[k]["work"] / (kinds[k]["ms"] * 1e-3) / 1e9
            breakdown[k]["frac_hbm"] = breakdown[k]["GBps"] / pk["hbm"]
    if "attention" in kinds:
        breakdown["attention"]["TFLOPs"] = kinds["attention"]["work"] / (kinds["attention"]["ms"] * 1e-3) / 1e12
    top = sorted(((v["ms"], k) for k, v in by_tag.items() if k[0] == "conv3d"), reverse=True)[:4]
    roof["top_shapes"] = [{"shape": k[1], "ms": by_tag[k]["ms"] / n_prof,
                           "TFLOPs": by_tag[k]["work"] / (by_tag[k]["ms"] * 1e-3) / 1e12} for _, k in top]

    # ---- second half of BASELINE.json's metric: generated volumes/s (full sampling + decode), one warm-up batch
    sampling = None
    if not args.no_sampling:
        del loop
        torch.cuda.empty_cache()
        ms_batch, _, _ = run_sampling(rank, world, local, args.sampling_T, 20, 1, 1)
        sampling = sampling_summary(ms_batch, world, args.sampling_T, 20)

    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_oracle_unet_step(os.cpu_count() or 1, budget_s=30.0, n_timed=1, n_warm=1)
    step_tf = FLOP_PER_VOLUME_STEP_16G * BATCH / (ms / args.steps * 1e-3) / 1e12
    line = {"metric": "unet_denoise_volume_steps_per_sec", "value": value, "unit": "volume-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "unet_step_16g: config_train_16g UNet (256,256,512) denoise step + fused DDPM update",
                       "batch_per_gpu": BATCH, "latent": list(LATENT_16G), "parallelism": f"replica-per-gpu x{world}",
                       "l2": "inputs larger than L2 (weights 0.3 GB + 181 MB activations per layer vs 126 MB L2)",
                       "cuda_graph": True, "batch_streams": len(loop_subs)},
            "step_tflops_algorithmic": step_tf, "frac_of_sustained_peak": step_tf / pk["bf16_sustained"],
            "e2e": {"value": e2e_value, "unit": "volume-steps/s", "h2d_bytes_per_step": nbytes,
                    "d2h_bytes_per_step": nbytes, "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps},
            "gpu_launches": per_graph * args.steps, "launches_per_step": per_graph,
            "roofline": roof, "kernel_breakdown": breakdown, "sampling": sampling, "cpu_baseline": cpu,
            "clocks": clk.summary()}
    print(json.dumps(line))


FLOP_UNET_32G = {16: 465.6e9, 20: 914.5e9, 40: 7910.7e9}     # per volume per step
FLOP_DEC_32G = {16: 888.7e9, 20: 1799.7e9, 40: 21738e9}      # per volume


def run_sampling(rank, world, local, T, latent, reps, warm):
    """BASELINE.json configs[2] per GPU: full DDPM latent sampling (config_train_32g UNet, concat conditioning) +
    AutoencoderKL decode, batch 8 volumes per GPU, final NCCL gather of volumes and seeds.  Returns (ms per batch
    [max over ranks], launches, clocks)."""
    import b200ldm
    from b200ldm import dist as bd, ops
    dev = torch.device("cuda", local)
    unet = seeded_init(b200ldm.DiffusionModelUNet(**UNET_32G, precision="bf16"), 1234).to(dev).eval()
    ae = seeded_init(b200ldm.AutoencoderKL(**AE_32G, precision="bf16"), 4321).to(dev).eval()
    sched = b200ldm.DDPMScheduler(**SCHED)
    sched.set_timesteps(T)
    inf = b200ldm.LatentDiffusionInferer(sched, scale_factor=1.0)
    lat = (16, latent, latent, latent)
    total = BATCH * world
    cond = torch.randn((total, 16) + lat[1:], generator=torch.Generator().manual_seed(7))
    seed = [42]

    def one():
        seed[0] += 1   # a new batch of volumes every call (new x_T, new in-kernel noise stream)
        return bd.sample_sharded(inf, ae, unet, sched, total, lat, base_seed=seed[0], conditioning=cond, mode="concat",
                                 device=dev)

    for _ in range(warm):
        one()
    l0 = ops.launch_count()
    with ClockSampler(local) as clk:
        ms = timed_region(one, reps, world)
    return ms / reps, (ops.launch_count() - l0) // reps, clk.summary()


def sampling_summary(ms_batch, world, T, latent):
    pk = peaks()
    vps = world * BATCH / (ms_batch * 1e-3)
    out = {"metric": "generated_volumes_per_sec", "value": vps, "unit": "volumes/s", "ms_per_batch": ms_batch,
           "workload": f"sample_32g: DDPM {T}-step latent sampling + AutoencoderKL decode, latent {latent}^3 -> "
                       f"{4 * latent}^3, concat conditioning, batch {BATCH}/GPU, final gather of volumes + seeds"}
    if latent in FLOP_UNET_32G:
        flop = BATCH * (T * FLOP_UNET_32G[latent] + FLOP_DEC_32G[latent])
        out["tflops_algorithmic"] = flop / (ms_batch * 1e-3) / 1e12
        out["frac_of_sustained_peak"] = out["tflops_algorithmic"] / pk["bf16_sustained"]
    return out


def bench_sample(args, rank, world, local):
    ms_batch, launches, clocks = run_sampling(rank, world, local, args.T, args.latent, args.steps, max(1, args.warmup))
    if rank != 0:
        return
    sm = sampling_summary(ms_batch, world, args.T, args.latent)
    line = {"metric": sm["metric"], "value": sm["value"], "unit": sm["unit"], "n_gpus": world, "steps": args.steps,
            "warmup": max(1, args.warmup), "ms_per_step": ms_batch, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": sm["workload"], "batch_per_gpu": BATCH, "parallelism": f"batch-sharded x{world}"},
            "tflops_algorithmic": sm.get("tflops_algorithmic"), "frac_of_sustained_peak": sm.get("frac_of_sustained_peak"),
            "gpu_launches": launches * args.steps, "clocks": clocks}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="unet_step_16g", choices=["unet_step_16g", "sample_32g"])
    ap.add_argument("--T", type=int, default=50, help="sample_32g: number of DDPM steps")
    ap.add_argument("--latent", type=int, default=20, help="sample_32g: latent edge")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sampling", action="store_true", help="skip the volumes/s leg of the default workload")
    ap.add_argument("--sampling-T", type=int, default=1000, help="DDPM steps of the volumes/s leg (reference: 1000)")
    args = ap.parse_args()
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
