"""Seeded full-size cases for the BASELINE.json configurations (cfg 3: DDPM-1000 at a 20^3 latent, cfg 4: DDIM-50 at
40^3, cfg 5: AutoencoderKL reconstruct at 96^3 ...), shared by scripts/long_parity.py (which runs the CPU oracle and
writes the fixtures / logs) and tests/test_gpu_long.py (which re-checks the CUDA path against those fixtures).

Model kwargs = 3d_ldm/config/config_train_32g.json:7-28,40-49; scheduler = 3d_ldm/inference.py:79-84; the sampling
call = 3d_ldm/train_diffusion.py:326-333 (conditioning latent, mode="concat")."""
from __future__ import annotations

import time

import torch

import oracle

UNET_32G = dict(spatial_dims=3, in_channels=32, out_channels=16, channels=[256, 512, 1024],
                attention_levels=[False, True, True], num_head_channels=[0, 64, 64], num_res_blocks=2)
AE_32G = dict(spatial_dims=3, in_channels=1, out_channels=1, latent_channels=16, channels=[64, 128, 256],
              num_res_blocks=2, norm_num_groups=32, norm_eps=1e-6, attention_levels=[False, False, True],
              with_encoder_nonlocal_attn=False, with_decoder_nonlocal_attn=False)
SCHED = dict(num_train_timesteps=1000, schedule="scaled_linear_beta", beta_start=0.0015, beta_end=0.0195)
DEV = "cuda"


def gen(seed):
    return torch.Generator(device="cpu").manual_seed(seed)


def psnr(a, b):
    """10 log10(peak^2 / mse), peak = max(ref) - min(ref) (SURVEY.md 8d parity gates)"""
    a, b = a.double().cpu(), b.double().cpu()
    peak = (b.max() - b.min()).item()
    mse = ((a - b) ** 2).mean().item()
    return 10.0 * torch.log10(torch.tensor(peak * peak / max(mse, 1e-30))).item()


def relerr(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def oracle_models():
    unet = oracle.randomize_parameters(oracle.DiffusionModelUNet(**UNET_32G), 1234).eval()
    ae = oracle.randomize_parameters(oracle.AutoencoderKL(**AE_32G), 4321).eval()
    return unet, ae


def port(ref_module, cls, kw, precision):
    m = cls(**kw, precision=precision)
    m.load_state_dict(ref_module.state_dict())
    return m.to(DEV).eval()


def sampling_inputs(latent: int, n_noise: int, batch: int = 1, seed: int = 42):
    """x_T, conditioning latent, per-step noise [n_noise, batch, 16, s, s, s] (n_noise = 0: none, deterministic DDIM)"""
    g = gen(seed)
    shp = (batch, 16, latent, latent, latent)
    x_T = torch.randn(shp, generator=g)
    cond = torch.randn(shp, generator=g)
    noise = torch.randn((n_noise,) + shp, generator=g) if n_noise else None
    return x_T, cond, noise


def scheduler(mod, kind: str, n_steps: int):
    cls = mod.DDPMScheduler if kind == "ddpm" else mod.DDIMScheduler
    s = cls(**SCHED)
    s.set_timesteps(n_steps)
    return s


def saved_steps(sched, every: int):
    """1-based step counts after which DiffusionInferer.sample(save_intermediates=True, intermediate_steps=every)
    stores the latent (t % every == 0)"""
    return [i + 1 for i, t in enumerate(sched.timesteps) if int(t) % every == 0]


def sample_gpu(unet, kind, n_steps, x_T, cond, noise, every):
    """latents at the saved steps + the final latent from the CUDA path (base-class sample: no decode per checkpoint)"""
    import b200ldm
    s = scheduler(b200ldm, kind, n_steps)
    inf = b200ldm.DiffusionInferer(s)
    lat, inter = inf.sample(input_noise=x_T.to(DEV), diffusion_model=unet, scheduler=s, save_intermediates=True,
                            intermediate_steps=every, conditioning=cond.to(DEV), mode="concat", verbose=False,
                            step_noise=None if noise is None else noise.to(DEV))
    return lat, inter


def run_sampling_vs_oracle(kind: str, latent: int, n_steps: int, every: int = 10, fixture: str | None = None):
    import b200ldm
    r_unet, r_ae = oracle_models()
    x_T, cond, noise = sampling_inputs(latent, n_steps if kind == "ddpm" else 0)
    # --- CUDA path (bf16 tensor-core engine)
    unet = port(r_unet, b200ldm.DiffusionModelUNet, UNET_32G, "bf16")
    ae = port(r_ae, b200ldm.AutoencoderKL, AE_32G, "bf16")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lat, inter = sample_gpu(unet, kind, n_steps, x_T, cond, noise, every)
    img = ae.decode_stage_2_outputs(lat)
    torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    # --- CPU oracle (fp32)
    s_ref = scheduler(oracle, kind, n_steps)
    t0 = time.perf_counter()
    o_lat, o_inter = oracle.DiffusionInferer(s_ref).sample(x_T, r_unet, s_ref, save_intermediates=True,
                                                           intermediate_steps=every, conditioning=cond, mode="concat",
                                                           verbose=False, step_noise=noise)
    t_unet = time.perf_counter() - t0
    with torch.no_grad():
        o_img = r_ae.decode_stage_2_outputs(o_lat)
    t_cpu = time.perf_counter() - t0
    steps = saved_steps(s_ref, every)
    assert len(steps) == len(inter) == len(o_inter), (len(steps), len(inter), len(o_inter))
    traj = {st: round(psnr(a, b), 2) for st, a, b in zip(steps, inter, o_inter)}
    # decode both final latents with the SAME (oracle) decoder as well: separates sampling error from decoder error
    with torch.no_grad():
        x_img = r_ae.decode_stage_2_outputs(lat.cpu())
    res = {
        "case": f"{kind}-{n_steps} @ {latent}^3 latent -> {4 * latent}^3, 32g UNet + AutoencoderKL, batch 1, bf16 vs CPU oracle",
        "psnr_latent_final_db": round(psnr(lat, o_lat), 2), "psnr_image_db": round(psnr(img, o_img), 2),
        "psnr_image_oracle_decoder_db": round(psnr(x_img, o_img), 2),
        "relerr_latent_final": relerr(lat, o_lat), "psnr_latent_by_step_db": traj,
        "frac_clamped_final": float((o_lat.abs() >= 1.0).double().mean()),
        "oracle_latent_std": float(o_lat.std()), "seconds_gpu_incl_capture": round(t_gpu, 2),
        "seconds_cpu_oracle": round(t_cpu, 1), "seconds_cpu_oracle_per_step": round(t_unet / n_steps, 3),
    }
    if fixture:
        keep = [s for s in (10, 100, 500, n_steps) if s in steps]
        torch.save({"kind": kind, "latent": latent, "n_steps": n_steps, "every": every,
                    "latents": {s: o_inter[steps.index(s)].clone() for s in keep}, "image": o_img.clone()}, fixture)
    return res


def run_sampling_bf16_vs_fp32(kind: str, latent: int, n_steps: int, batch: int, every: int = 10):
    """whole batch: the bf16 engine against the repo's own fp32 validation engine (itself <= 1e-4 per step vs the oracle)"""
    import b200ldm
    r_unet, r_ae = oracle_models()
    x_T, cond, noise = sampling_inputs(latent, n_steps if kind == "ddpm" else 0, batch)
    out = {}
    for prec in ("bf16", "fp32"):
        unet = port(r_unet, b200ldm.DiffusionModelUNet, UNET_32G, prec)
        ae = port(r_ae, b200ldm.AutoencoderKL, AE_32G, prec)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lat, inter = sample_gpu(unet, kind, n_steps, x_T, cond, noise, every)
        img = ae.decode_stage_2_outputs(lat)
        torch.cuda.synchronize()
        out[prec] = (lat.cpu(), [i.cpu() for i in inter], img.cpu(), time.perf_counter() - t0)
        del unet, ae
        torch.cuda.empty_cache()
    s_ref = scheduler(oracle, kind, n_steps)
    steps = saved_steps(s_ref, every)
    a, b = out["bf16"], out["fp32"]
    per_vol = [round(psnr(a[2][i], b[2][i]), 2) for i in range(batch)]
    return {
        "case": f"{kind}-{n_steps} @ {latent}^3 latent, batch {batch}, bf16 engine vs fp32 validation engine",
        "psnr_image_db": round(psnr(a[2], b[2]), 2), "psnr_image_per_volume_db": per_vol,
        "psnr_latent_final_db": round(psnr(a[0], b[0]), 2),
        "psnr_latent_by_step_db": {st: round(psnr(x, y), 2) for st, x, y in zip(steps, a[1], b[1])},
        "seconds_bf16": round(a[3], 2), "seconds_fp32": round(b[3], 2),
    }


def run_cfg5(edge: int):
    """AutoencoderKL reconstruct = decode(encode(x).mu) on a uniform[0,1) volume (3d_ldm/utils.py:97 scales to [0,1])"""
    import b200ldm
    _, r_ae = oracle_models()
    x = torch.rand((1, 1, edge, edge, edge), generator=gen(9))
    ae = port(r_ae, b200ldm.AutoencoderKL, AE_32G, "bf16")
    mu, sigma = ae.encode(x.to(DEV))
    rec = ae.reconstruct(x.to(DEV))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        o_mu, o_sigma = r_ae.encode(x)
        o_rec = r_ae.decode(o_mu)
        # decoder alone on the oracle's latent: separates the two halves
        rec_from_o = ae.decode(o_mu.to(DEV))
    t_cpu = time.perf_counter() - t0
    return {
        "case": f"AutoencoderKL (32g) reconstruct {edge}^3, batch 1, bf16 vs CPU oracle",
        "relerr_mu": relerr(mu, o_mu), "psnr_mu_db": round(psnr(mu, o_mu), 2), "relerr_sigma": relerr(sigma, o_sigma),
        "psnr_reconstruct_db": round(psnr(rec, o_rec), 2), "relerr_reconstruct": relerr(rec, o_rec),
        "psnr_decode_of_oracle_latent_db": round(psnr(rec_from_o, o_rec), 2),
        "seconds_cpu_oracle": round(t_cpu, 1),
    }
