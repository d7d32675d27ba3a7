"""GPU parity of the drop-in modules against the oracle and the committed golden vectors:
per-step noise prediction (<= 1e-4 fp32 validation mode, <= 2e-2 bf16), AutoencoderKL encode / decode,
and the full LatentDiffusionInferer.sample loop (decoded volume PSNR >= 40 dB) - BASELINE.md 6."""
import os

import pytest
import torch

import golden_cases as gc
import oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(__file__), "golden", "tiny_cases.pt")
EPS_TOL = {"fp32": 1e-4, "bf16": 2e-2}


def relerr(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def psnr(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    peak = (b.max() - b.min()).item()
    mse = ((a - b) ** 2).mean().item()
    return 10.0 * torch.log10(torch.tensor(peak * peak / max(mse, 1e-30))).item()


def port(ref_module, cls, kw, precision):
    m = cls(**kw, precision=precision)
    m.load_state_dict(ref_module.state_dict())
    return m.to(DEV).eval()


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=True)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_unet_eps_vs_oracle_and_golden(precision, gold):
    import b200ldm
    ref, x, ts = gc.tiny_unet_case()
    m = port(ref, b200ldm.DiffusionModelUNet, gc.UNET_KW, precision)
    got = m(x.to(DEV), ts.to(DEV))
    with torch.no_grad():
        want = ref(x, ts)
    assert relerr(want, gold["unet_eps"]) <= 1e-5          # oracle still equals its committed fixture
    e = relerr(got, gold["unet_eps"])
    assert e <= EPS_TOL[precision], e
    # single timestep broadcast over the batch (what inferer.sample passes: Tensor((t,)))
    got1 = m(x.to(DEV), ts[:1].to(DEV))
    with torch.no_grad():
        want1 = ref(x, ts[:1])
    assert relerr(got1, want1) <= EPS_TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_autoencoder_decode_encode_vs_golden(precision, gold):
    import b200ldm
    ref, z = gc.tiny_ae_case()
    m = port(ref, b200ldm.AutoencoderKL, gc.AE_KW, precision)
    got = m.decode_stage_2_outputs(z.to(DEV))
    tol = 1e-4 if precision == "fp32" else 2e-2
    assert relerr(got, gold["ae_decode"]) <= tol, relerr(got, gold["ae_decode"])
    mu, sigma = m.encode(gc.tiny_ae_image().to(DEV))
    assert relerr(mu, gold["ae_mu"]) <= tol, relerr(mu, gold["ae_mu"])
    assert relerr(sigma, gold["ae_sigma"]) <= (1e-3 if precision == "fp32" else 1e-1)  # exp() amplifies
    nz = torch.randn(mu.shape, generator=gc.gen(5))
    zs = m.sampling(mu, sigma, noise=nz.to(DEV))
    assert relerr(zs, mu.cpu() + sigma.cpu() * nz) <= 1e-6
    rec = m.reconstruct(gc.tiny_ae_image().to(DEV))
    with torch.no_grad():
        assert relerr(rec, ref.reconstruct(gc.tiny_ae_image())) <= (2e-4 if precision == "fp32" else 4e-2)


@pytest.mark.parametrize("kind", ["ddpm", "ddim"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_latent_sampling_vs_golden(kind, precision, gold):
    """inferer.sample(..., conditioning, mode="concat") as at 3d_ldm/train_diffusion.py:326-333, 5 steps,
    identical weights / x_T / per-step noise on both sides."""
    import b200ldm
    unet = port(gc.tiny_unet(), b200ldm.DiffusionModelUNet, gc.UNET_KW, precision)
    ae = port(gc.tiny_ae(), b200ldm.AutoencoderKL, gc.AE_KW, precision)
    sched = gc.make_scheduler(b200ldm, kind, 5)
    x_T, cond, step_noise = gc.sampling_inputs(5)
    inf = b200ldm.LatentDiffusionInferer(sched, scale_factor=1.0)
    img, lat = inf.sample(input_noise=x_T.to(DEV), autoencoder_model=ae, diffusion_model=unet, scheduler=sched,
                          conditioning=cond.to(DEV), mode="concat", verbose=False, step_noise=step_noise.to(DEV),
                          return_latent=True)
    p_lat, p_img = psnr(lat, gold[f"sample_{kind}_latent"]), psnr(img, gold[f"sample_{kind}_image"])
    print(f"{kind}/{precision}: latent PSNR {p_lat:.1f} dB, image PSNR {p_img:.1f} dB")
    if precision == "fp32":
        assert relerr(lat, gold[f"sample_{kind}_latent"]) <= 1e-3
        assert relerr(img, gold[f"sample_{kind}_image"]) <= 1e-3
    assert p_img >= 40.0 and p_lat >= 35.0, (p_lat, p_img)  # the contract is on the decoded volume


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sampling_graph_equals_eager_and_intermediates(precision):
    import b200ldm
    from b200ldm import inferer as inf_mod
    # GroupNorm partial sums use atomics, so runs are not bit-identical; in bf16 a 1-ulp flip is 4e-3
    tol = 1e-4 if precision == "fp32" else 2e-2
    unet = port(gc.tiny_unet(), b200ldm.DiffusionModelUNet, gc.UNET_KW, precision)
    ae = port(gc.tiny_ae(), b200ldm.AutoencoderKL, gc.AE_KW, precision)
    sched = gc.make_scheduler(b200ldm, "ddpm", 5)
    x_T, cond, step_noise = gc.sampling_inputs(5)
    inf = b200ldm.LatentDiffusionInferer(sched)
    kw = dict(input_noise=x_T.to(DEV), autoencoder_model=ae, diffusion_model=unet, scheduler=sched,
              conditioning=cond.to(DEV), mode="concat", verbose=False, step_noise=step_noise.to(DEV))
    a = inf.sample(**kw)
    old = inf_mod.USE_GRAPH
    inf_mod.USE_GRAPH = False
    try:
        b = inf.sample(**kw)
    finally:
        inf_mod.USE_GRAPH = old
    assert relerr(a, b) <= tol, relerr(a, b)
    img, inter = inf.sample(**kw, save_intermediates=True, intermediate_steps=400)
    assert len(inter) == sum(1 for t in sched.timesteps if int(t) % 400 == 0)
    assert inter[-1].shape == img.shape
    # in-kernel Philox noise: deterministic in the seed
    kw.pop("step_noise")
    c1, c2, c3 = inf.sample(**kw, seed=7), inf.sample(**kw, seed=7), inf.sample(**kw, seed=8)
    assert relerr(c1, c2) <= tol and relerr(c1, c3) > 5e-2


def test_cached_sampling_loop_tracks_inputs_and_weights():
    """sample() keeps one captured loop per (model, shape, schedule): a second call must see the NEW noise /
    conditioning / seed, and rewritten weights must invalidate the captured graphs (UNet step and decode)."""
    import b200ldm
    unet = port(gc.tiny_unet(), b200ldm.DiffusionModelUNet, gc.UNET_KW, "bf16")
    ae = port(gc.tiny_ae(), b200ldm.AutoencoderKL, gc.AE_KW, "bf16")
    sched = gc.make_scheduler(b200ldm, "ddpm", 5)
    x_T, cond, step_noise = gc.sampling_inputs(5)
    g = torch.Generator().manual_seed(99)
    x2, c2 = torch.randn(x_T.shape, generator=g), torch.randn(cond.shape, generator=g)
    inf = b200ldm.LatentDiffusionInferer(sched)

    def run(i, x, c):
        return i.sample(input_noise=x.to(DEV), autoencoder_model=ae, diffusion_model=unet, scheduler=sched,
                        conditioning=c.to(DEV), mode="concat", verbose=False, step_noise=step_noise.to(DEV))

    a1 = run(inf, x_T, cond)
    b1 = run(inf, x2, c2)                                   # cached loop, new inputs
    assert len(inf._loops) == 1
    b_fresh = run(b200ldm.LatentDiffusionInferer(sched), x2, c2)
    assert relerr(b1, b_fresh) <= 2e-2 and relerr(b1, a1) > 5e-2
    a2 = run(inf, x_T, cond)
    assert relerr(a2, a1) <= 2e-2
    # new weights: both cached graphs hold stale packed-weight pointers and must be rebuilt
    sd = {k: v * 1.25 for k, v in unet.state_dict().items()}
    unet.load_state_dict(sd)
    ae.load_state_dict({k: v * 0.9 for k, v in ae.state_dict().items()})
    a3 = run(inf, x_T, cond)
    a3_fresh = run(b200ldm.LatentDiffusionInferer(sched), x_T, cond)
    assert relerr(a3, a3_fresh) <= 2e-2 and relerr(a3, a1) > 5e-2


def test_serving_generator_batches_requests():
    """LatentDiffusionGenerator.generate == the reference request loop's `model.generate(noise, ...)`, batched: results
    are seed-deterministic, shaped like the request, finite, and padded batches reuse one captured loop."""
    import b200ldm
    unet = port(gc.tiny_unet(), b200ldm.DiffusionModelUNet, gc.UNET_KW, "bf16")
    ae = port(gc.tiny_ae(), b200ldm.AutoencoderKL, gc.AE_KW, "bf16")
    gen = b200ldm.LatentDiffusionGenerator(ae, unet, dict(num_train_timesteps=1000, schedule="scaled_linear_beta",
                                                          beta_start=0.0015, beta_end=0.0195), sampler="ddpm", max_batch=2,
                                           downsample=2)       # the tiny AE has one down-sampling level
    patch = tuple(2 * v for v in gc.LATENT)
    a = gen.generate_many(3, patch, DEV, num_inference_steps=4, seed=11)
    b = gen.generate_many(3, patch, DEV, num_inference_steps=4, seed=11)
    c = gen.generate_many(3, patch, DEV, num_inference_steps=4, seed=12)
    assert a.shape == (3, 1) + patch and bool(torch.isfinite(a).all())
    assert relerr(a, b) <= 2e-2 and relerr(a, c) > 5e-2
    assert len(gen.inferer._loops) == 1                      # 2 + 1(padded to 2): one shape, one captured graph
    out = gen(torch.empty((1, 1) + patch, device=DEV))       # `model(noise)` branch
    assert out.shape == (1, 1) + patch


def test_unseeded_sample_calls_draw_fresh_noise():
    """without `seed=` every sample() call must use a new per-step noise stream (the reference draws
    fresh torch.randn at every step of every call, 3d_ldm/inference.py:91-99); torch.manual_seed still makes the whole
    sequence reproducible."""
    import b200ldm
    unet = port(gc.tiny_unet(), b200ldm.DiffusionModelUNet, gc.UNET_KW, "bf16")
    ae = port(gc.tiny_ae(), b200ldm.AutoencoderKL, gc.AE_KW, "bf16")
    sched = gc.make_scheduler(b200ldm, "ddpm", 5)
    x_T, cond, _ = gc.sampling_inputs(5)
    inf = b200ldm.LatentDiffusionInferer(sched)
    kw = dict(input_noise=x_T.to(DEV), autoencoder_model=ae, diffusion_model=unet, scheduler=sched,
              conditioning=cond.to(DEV), mode="concat", verbose=False, return_latent=True)
    torch.manual_seed(42)
    _, a = inf.sample(**kw)
    _, b = inf.sample(**kw)          # same x_T, same cached loop, no seed: must differ
    assert relerr(a, b) > 5e-2, relerr(a, b)
    torch.manual_seed(42)
    _, a2 = inf.sample(**kw)
    _, b2 = inf.sample(**kw)
    assert relerr(a2, a) <= 2e-2 and relerr(b2, b) <= 2e-2
    _, c = inf.sample(**kw, seed=7)
    _, c2 = inf.sample(**kw, seed=7)
    assert relerr(c, c2) <= 2e-2


@pytest.mark.parametrize("kind", ["ddpm", "ddim"])
def test_scheduler_step_honours_generator(kind):
    """scheduler.step(..., generator=g) draws this step's noise from g with torch.randn exactly as upstream does, so the
    same generator state gives the oracle's result (VERDICT r01 missing #5)."""
    import b200ldm
    ours, ref = gc.make_scheduler(b200ldm, kind, 10), gc.make_scheduler(oracle, kind, 10)
    g0 = gc.gen(3)
    x = torch.randn((2, 4) + gc.LATENT, generator=g0)
    eps = torch.randn((2, 4) + gc.LATENT, generator=g0)
    t = int(ours.timesteps[3])
    extra = {"eta": 0.7} if kind == "ddim" else {}
    want, want0 = ref.step(eps, t, x, generator=gc.gen(99), **extra)
    got, got0 = ours.step(eps.to(DEV), t, x.to(DEV), generator=gc.gen(99), **extra)
    assert relerr(got, want) <= 1e-5 and relerr(got0, want0) <= 1e-5
    other, _ = ours.step(eps.to(DEV), t, x.to(DEV), generator=gc.gen(100), **extra)
    assert relerr(other, want) > 1e-3
    gd = torch.Generator(device=DEV).manual_seed(5)      # a generator that lives on the GPU, as upstream requires there
    a, _ = ours.step(eps.to(DEV), t, x.to(DEV), generator=gd, **extra)
    gd.manual_seed(5)
    b, _ = ours.step(eps.to(DEV), t, x.to(DEV), generator=gd, **extra)
    assert torch.equal(a, b)


def test_sampling_fused_scheduler_equals_separate_kernel():
    """north_star (4): the scheduler update rides in the UNet's last kernel.  Same trajectory as with the separate
    scheduler kernel (which reads a bf16-rounded prediction instead of the fp32 sums)."""
    import b200ldm
    from b200ldm import inferer as inf_mod
    unet = port(gc.tiny_unet(), b200ldm.DiffusionModelUNet, gc.UNET_KW, "bf16")
    ae = port(gc.tiny_ae(), b200ldm.AutoencoderKL, gc.AE_KW, "bf16")
    sched = gc.make_scheduler(b200ldm, "ddpm", 5)
    x_T, cond, step_noise = gc.sampling_inputs(5)
    kw = dict(input_noise=x_T.to(DEV), autoencoder_model=ae, diffusion_model=unet, scheduler=sched,
              conditioning=cond.to(DEV), mode="concat", verbose=False, step_noise=step_noise.to(DEV), return_latent=True)
    assert inf_mod.FUSE_SCHED and unet.out[2].conv.tap_expanded(torch.empty(0, dtype=torch.bfloat16))
    inf = b200ldm.LatentDiffusionInferer(sched)
    img_f, lat_f = inf.sample(**kw)
    assert inf._loops and any(sp is not None for sp in next(iter(inf._loops.values()))._sched)      # the fused path really ran
    inf_mod.FUSE_SCHED = False
    try:
        img_s, lat_s = b200ldm.LatentDiffusionInferer(sched).sample(**kw)
    finally:
        inf_mod.FUSE_SCHED = True
    assert relerr(lat_f, lat_s) <= 1e-2 and relerr(img_f, img_s) <= 2e-2


def test_inferer_call_matches_oracle():
    """LatentDiffusionInferer.__call__ (3d_ldm/train_diffusion.py:197-205): encode -> add_noise -> concat -> UNet."""
    import b200ldm
    r_unet, r_ae = gc.tiny_unet(), gc.tiny_ae()
    unet = port(r_unet, b200ldm.DiffusionModelUNet, gc.UNET_KW, "fp32")
    ae = port(r_ae, b200ldm.AutoencoderKL, gc.AE_KW, "fp32")
    img = torch.rand((2, 1) + tuple(2 * v for v in gc.LATENT), generator=gc.gen(77))
    cond = torch.randn((2, 4) + gc.LATENT, generator=gc.gen(78))
    noise = torch.randn((2, 4) + gc.LATENT, generator=gc.gen(79))
    t = torch.tensor([10, 900])
    ours = b200ldm.LatentDiffusionInferer(gc.make_scheduler(b200ldm, "ddpm", 1000), scale_factor=0.8)
    ref = oracle.LatentDiffusionInferer(gc.make_scheduler(oracle, "ddpm", 1000), scale_factor=0.8)
    zero = torch.zeros((2, 4) + gc.LATENT)
    got = ours(inputs=img.to(DEV), autoencoder_model=ae, diffusion_model=unet, noise=noise.to(DEV),
               timesteps=t.to(DEV), condition=cond.to(DEV), mode="concat", encode_noise=zero.to(DEV))
    with torch.no_grad():
        mu, _ = r_ae.encode(img)
        noisy = ref.scheduler.add_noise(mu * 0.8, noise, t)
        want = r_unet(torch.cat([noisy, cond], 1), timesteps=t)
    assert relerr(got, want) <= 2e-4, relerr(got, want)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_unet_32g_step_vs_oracle(precision):
    """The reference's default UNet (3d_ldm/config/config_train_32g.json:40-49), 16^3 latent, one step."""
    import b200ldm
    kw = dict(spatial_dims=3, in_channels=32, out_channels=16, channels=[256, 512, 1024],
              attention_levels=[False, True, True], num_head_channels=[0, 64, 64], num_res_blocks=2)
    torch.manual_seed(0)
    ref = oracle.randomize_parameters(oracle.DiffusionModelUNet(**kw), 1234).eval()
    x = torch.randn((1, 32, 16, 16, 16), generator=gc.gen(3))
    ts = torch.tensor([421.0])
    m = port(ref, b200ldm.DiffusionModelUNet, kw, precision)
    got = m(x.to(DEV), ts.to(DEV))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        want = ref(x, ts)
    e = relerr(got, want)
    print(f"unet32g/{precision}: eps relerr {e:.3e}")
    assert e <= EPS_TOL[precision], e


def test_unet_16g_cfg2_full_size_parity_and_batch_independence():
    """BASELINE.json configs[1] at its FULL size (config_train_16g UNet, latent 36x44x28, bf16): one volume against the CPU
    oracle (per-step eps <= 2e-2), then the size-independent property the domain offers for the full batch of 8 -
    volumes are independent, so row i of the batched prediction must equal the batch-1 prediction of volume i."""
    import b200ldm
    kw = dict(spatial_dims=3, in_channels=8, out_channels=8, channels=[256, 256, 512],
              attention_levels=[False, True, True], num_head_channels=[0, 64, 64], num_res_blocks=2)
    ref = oracle.randomize_parameters(oracle.DiffusionModelUNet(**kw), 1234).eval()
    m = port(ref, b200ldm.DiffusionModelUNet, kw, "bf16")
    x = torch.randn((8, 8, 36, 44, 28), generator=gc.gen(5))
    ts = torch.tensor([637.0])
    got8 = m(x.to(DEV), ts.to(DEV))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        want0 = ref(x[:1], ts)
    e = relerr(got8[:1], want0)
    print(f"unet16g cfg2 full size / bf16: eps relerr {e:.3e}")
    assert e <= EPS_TOL["bf16"], e
    for i in (3, 7):
        gi = m(x[i:i + 1].to(DEV), ts.to(DEV))
        d = relerr(got8[i:i + 1], gi)
        # other tile plans / split factors / atomics order for batch 1: two bf16 evaluations of the same function, each
        # within ~1e-2 of the fp32 answer, so their distance is bounded by the bf16 tolerance itself (measured 0.6-1.1e-2)
        assert d <= EPS_TOL["bf16"], (i, d)


def test_inference_cli_dropin(tmp_path):
    """The reference CLI flow (3d_ldm/inference.py:31-102) through b200ldm: JSON config with MONAI `_target_`s ->
    define_instance -> DDPMScheduler -> LatentDiffusionInferer.sample -> NIfTI on disk."""
    import json
    import subprocess
    import sys
    from b200ldm.nifti import load_nifti
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = {
        "spatial_dims": 3, "image_channels": 1, "latent_channels": 4,
        "autoencoder_def": {"_target_": "monai.networks.nets.AutoencoderKL", "spatial_dims": "@spatial_dims",
                            "in_channels": "$@image_channels", "out_channels": "@image_channels",
                            "latent_channels": "@latent_channels", "channels": [64, 128, 128], "num_res_blocks": 1,
                            "norm_num_groups": 32, "norm_eps": 1e-06, "attention_levels": [False, False, True],
                            "with_encoder_nonlocal_attn": False, "with_decoder_nonlocal_attn": False},
        "diffusion_def": {"_target_": "monai.networks.nets.DiffusionModelUNet", "spatial_dims": "@spatial_dims",
                          "in_channels": 8, "out_channels": "@latent_channels", "channels": [64, 128],
                          "attention_levels": [False, True], "num_head_channels": [0, 64], "num_res_blocks": 1},
        "diffusion_train": {"patch_size": [32, 48, 16]},
        "NoiseScheduler": {"num_train_timesteps": 1000, "beta_start": 0.0015, "beta_end": 0.0195},
    }
    env = {"model_dir": str(tmp_path), "output_dir": str(tmp_path / "out")}
    (tmp_path / "cfg.json").write_text(json.dumps(cfg))
    (tmp_path / "env.json").write_text(json.dumps(env))
    cli = os.path.join(root, "3d-latent-diffusion-model_b200", "inference.py")
    r = subprocess.run([sys.executable, cli, "-e", str(tmp_path / "env.json"), "-c", str(tmp_path / "cfg.json"), "-n", "2",
                        "--steps", "4", "--random-init"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    outs = sorted(os.listdir(tmp_path / "out"))
    assert len(outs) == 2 and all(o.endswith(".nii") for o in outs)
    vol, aff = load_nifti(str(tmp_path / "out" / outs[0]))
    assert vol.shape == (32, 48, 16, 1) and bool((vol == vol).all())


def test_cfg1_default_config_sampling_psnr():
    """BASELINE.json configs[0]: the reference's default config_train_32g UNet + AutoencoderKL, 1 x 16^3 latent -> 64^3
    volume, 10 DDPM steps (inference.py flow with concat conditioning), bf16 tensor-core path vs the fp32 CPU oracle on
    identical weights / x_T / per-step noise: decoded volume PSNR >= 40 dB."""
    import b200ldm
    ukw = dict(spatial_dims=3, in_channels=32, out_channels=16, channels=[256, 512, 1024],
               attention_levels=[False, True, True], num_head_channels=[0, 64, 64], num_res_blocks=2)
    akw = dict(spatial_dims=3, in_channels=1, out_channels=1, latent_channels=16, channels=[64, 128, 256],
               num_res_blocks=2, norm_num_groups=32, norm_eps=1e-6, attention_levels=[False, False, True],
               with_encoder_nonlocal_attn=False, with_decoder_nonlocal_attn=False)
    r_unet = oracle.randomize_parameters(oracle.DiffusionModelUNet(**ukw), 1234).eval()
    r_ae = oracle.randomize_parameters(oracle.AutoencoderKL(**akw), 4321).eval()
    unet = port(r_unet, b200ldm.DiffusionModelUNet, ukw, "bf16")
    ae = port(r_ae, b200ldm.AutoencoderKL, akw, "bf16")
    n = 10
    g = gc.gen(42)
    x_T = torch.randn((1, 16, 16, 16, 16), generator=g)
    cond = torch.randn((1, 16, 16, 16, 16), generator=g)
    step_noise = torch.randn((n, 1, 16, 16, 16, 16), generator=g)
    s_ours, s_ref = gc.make_scheduler(b200ldm, "ddpm", n), gc.make_scheduler(oracle, "ddpm", n)
    img = b200ldm.LatentDiffusionInferer(s_ours, 1.0).sample(
        input_noise=x_T.to(DEV), autoencoder_model=ae, diffusion_model=unet, scheduler=s_ours,
        conditioning=cond.to(DEV), mode="concat", verbose=False, step_noise=step_noise.to(DEV))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    want = oracle.LatentDiffusionInferer(s_ref, 1.0).sample(x_T, r_ae, r_unet, s_ref, conditioning=cond, mode="concat",
                                                            verbose=False, step_noise=step_noise)
    assert img.shape == (1, 1, 64, 64, 64)
    p = psnr(img, want)
    print(f"cfg1 (32g, 16^3 -> 64^3, DDPM-10, bf16): decoded volume PSNR {p:.1f} dB")
    assert p >= 40.0, p


# ------------------------------------------------------------------------------------------
# cross-attention conditioning (with_conditioning=True, SpatialTransformer)
# ------------------------------------------------------------------------------------------
XGOLD = os.path.join(os.path.dirname(__file__), "golden", "xattn_cases.pt")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_xattn_unet_eps_vs_oracle_and_golden(precision):
    import b200ldm
    xgold = torch.load(XGOLD, weights_only=True)
    ref, x, ts, ctx, labels = gc.tiny_xattn_case()
    m = port(ref, b200ldm.DiffusionModelUNet, gc.XATTN_UNET_KW, precision)
    args = (x.to(DEV), ts.to(DEV))
    got = m(*args, context=ctx.to(DEV), class_labels=labels.to(DEV))            # eager (first call of the signature)
    e = relerr(got, xgold["xattn_unet_eps"])
    assert e <= EPS_TOL[precision], e
    got2 = m(*args, context=ctx.to(DEV), class_labels=labels.to(DEV))           # captures the graph
    got3 = m(*args, context=ctx.flip(0).to(DEV), class_labels=labels.flip(0).to(DEV))   # replays it with new inputs
    assert relerr(got2, got) <= (1e-4 if precision == "fp32" else 2e-2)   # atomics / measured kernel choice: not bitwise
    with torch.no_grad():
        want3 = ref(x, ts, context=ctx.flip(0), class_labels=labels.flip(0))
    assert relerr(got3, want3) <= EPS_TOL[precision]
    assert relerr(got3, got) > 1e-3       # the context / labels really feed the prediction
    with pytest.raises(ValueError):
        m(*args, class_labels=labels.to(DEV))                                    # with_conditioning needs a context


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_xattn_latent_sampling_vs_golden(precision):
    """inferer.sample(..., conditioning=context, mode="crossattn") - the reference's commented-out call
    (3d_ldm/train_diffusion.py:314-321) - against the oracle's trajectory on the same noise."""
    import b200ldm
    xgold = torch.load(XGOLD, weights_only=True)
    (r_unet, r_ae, x_T, ctx, step_noise), _, _ = gc.tiny_xattn_sampling_case("ddpm")
    kw = {**gc.XATTN_UNET_KW, "num_class_embeds": None}
    unet = port(r_unet, b200ldm.DiffusionModelUNet, kw, precision)
    ae = port(r_ae, b200ldm.AutoencoderKL, gc.AE_KW, precision)
    sched = gc.make_scheduler(b200ldm, "ddpm", 4)
    inf = b200ldm.LatentDiffusionInferer(sched, 1.0)
    for rep in range(2):   # second call: cached loop + graph, context buffer refilled in place
        img, lat = inf.sample(x_T.to(DEV), ae, unet, sched, conditioning=ctx.to(DEV), mode="crossattn", verbose=False,
                              step_noise=step_noise.to(DEV), return_latent=True)
        assert relerr(lat, xgold["xattn_sample_latent"]) <= (2e-4 if precision == "fp32" else 5e-2)
        assert psnr(img, xgold["xattn_sample_image"]) >= (80.0 if precision == "fp32" else 40.0)
    plain = port(gc.tiny_unet(), b200ldm.DiffusionModelUNet, gc.UNET_KW, precision)
    with pytest.raises(ValueError):   # context handed to a model without cross attention
        inf.sample(x_T.to(DEV), ae, plain, sched, conditioning=ctx.to(DEV), mode="crossattn", verbose=False)


# ------------------------------------------------------------------------------------------
# learned variance (DDPMScheduler variance_type "learned" / "learned_range")
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("vt", ["learned", "learned_range"])
def test_scheduler_step_learned_variance_vs_oracle(vt):
    import b200ldm
    ours, ref = gc.make_scheduler(b200ldm, "ddpm", 10, variance_type=vt), gc.make_scheduler(oracle, "ddpm", 10, variance_type=vt)
    g = gc.gen(17)
    x, e, z = (torch.randn((2, 4) + gc.LATENT, generator=g) for _ in range(3))
    v = torch.rand((2, 4) + gc.LATENT, generator=g) * (1.6 if vt == "learned_range" else 0.02) \
        - (0.8 if vt == "learned_range" else 0.0)
    out = torch.cat([e, v], dim=1)
    for t in (900, 300, 0):
        got, got0 = ours.step(out.to(DEV), t, x.to(DEV), noise=z.to(DEV))
        want, want0 = ref.step(out, t, x, noise=z)
        assert relerr(got, want) <= 2e-6 and relerr(got0, want0) <= 2e-6, (t, relerr(got, want))
        plain, _ = ours.step(e.to(DEV), t, x.to(DEV), noise=z.to(DEV))     # falls back to the unclamped fixed variance
        assert relerr(plain, ref.step(e, t, x, noise=z)[0]) <= 2e-6
    with pytest.raises(ValueError):
        gc.make_scheduler(b200ldm, "ddpm", 10).step(out.to(DEV), 900, x.to(DEV))   # 2C channels need a learned type


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sampling_with_learned_range_variance(precision):
    """UNet with 2 x latent channels out + learned_range variance through inferer.sample (graph replay), vs the oracle."""
    import b200ldm
    kw = {**gc.UNET_KW, "out_channels": 8}
    r_unet = oracle.randomize_parameters(oracle.DiffusionModelUNet(**kw), 99).eval()
    r_ae = gc.tiny_ae()
    unet = port(r_unet, b200ldm.DiffusionModelUNet, kw, precision)
    ae = port(r_ae, b200ldm.AutoencoderKL, gc.AE_KW, precision)
    n = 4
    sched = gc.make_scheduler(b200ldm, "ddpm", n, variance_type="learned_range")
    r_sched = gc.make_scheduler(oracle, "ddpm", n, variance_type="learned_range")
    x_T, cond, step_noise = gc.sampling_inputs(n)
    img, lat = b200ldm.LatentDiffusionInferer(sched, 1.0).sample(
        x_T.to(DEV), ae, unet, sched, conditioning=cond.to(DEV), mode="concat", verbose=False,
        step_noise=step_noise.to(DEV), return_latent=True)
    want_img, want_lat = oracle.LatentDiffusionInferer(r_sched, 1.0).sample(
        x_T, r_ae, r_unet, r_sched, conditioning=cond, mode="concat", verbose=False, step_noise=step_noise,
        return_latent=True)
    assert relerr(lat, want_lat) <= (2e-4 if precision == "fp32" else 5e-2), relerr(lat, want_lat)
    assert psnr(img, want_img) >= (80.0 if precision == "fp32" else 40.0)
    fixed = gc.make_scheduler(b200ldm, "ddpm", n)
    with pytest.raises(ValueError):   # 2C-channel UNet with a fixed-variance scheduler
        b200ldm.LatentDiffusionInferer(fixed, 1.0).sample(x_T.to(DEV), ae, unet, fixed, conditioning=cond.to(DEV),
                                                          mode="concat", verbose=False)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sub_batch_streams_draw_the_same_noise(precision):
    """SamplingLoop(batch_streams=n): the batch runs as n sub-batches on parallel graph branches; in-kernel Philox noise is
    keyed by the element's index in the WHOLE batch (philox_offset), so the trajectory does not depend on the split."""
    import b200ldm
    from b200ldm.inferer import SamplingLoop
    unet = port(gc.tiny_unet(), b200ldm.DiffusionModelUNet, gc.UNET_KW, precision)
    sched = gc.make_scheduler(b200ldm, "ddpm", 4)
    g = gc.gen(77)
    x_T, cond = torch.randn((3, 4) + gc.LATENT, generator=g), torch.randn((3, 4) + gc.LATENT, generator=g)
    outs = []
    for n in (1, 2, 3):
        loop = SamplingLoop(unet, sched, x_T.to(DEV), cond.to(DEV), "concat", seed=1234, use_graph=True, batch_streams=n)
        assert len(loop.subs) == n and loop.subs[0][0] == 0 and loop.subs[-1][1] == 3
        for _ in sched.timesteps:
            loop.step()
        outs.append(loop.latent_ncdhw())
    tol = 1e-4 if precision == "fp32" else 2e-2
    assert relerr(outs[1], outs[0]) <= tol and relerr(outs[2], outs[0]) <= tol
    other = SamplingLoop(unet, sched, x_T.to(DEV), cond.to(DEV), "concat", seed=1235, use_graph=True, batch_streams=2)
    for _ in sched.timesteps:
        other.step()
    assert relerr(other.latent_ncdhw(), outs[0]) > 5e-2      # ... and really depends on the seed


def test_xattn_unet_realistic_widths_bf16():
    """cross-attention conditioning at realistic widths: 512-channel transformer level (8 heads of 64 -> the tcgen05
    attention kernels for both the 216-token self attention and the 77-token context), 768-wide context, GEGLU width
    2048, two-row LayerNorm chunks - bf16 engine against the fp32 oracle on the CPU."""
    import b200ldm
    kw = dict(spatial_dims=3, in_channels=8, out_channels=8, channels=(256, 512), attention_levels=(False, True),
              num_head_channels=(0, 64), num_res_blocks=1, norm_num_groups=32, norm_eps=1e-6, with_conditioning=True,
              cross_attention_dim=768, transformer_num_layers=1)
    ref = oracle.randomize_parameters(oracle.DiffusionModelUNet(**kw), 31).eval()
    g = gc.gen(32)
    x, ctx = torch.randn((2, 8, 12, 12, 12), generator=g), torch.randn((2, 77, 768), generator=g)
    ts = torch.tensor([250.0, 731.0])
    with torch.no_grad():
        want = ref(x, ts, context=ctx)
    m = port(ref, b200ldm.DiffusionModelUNet, kw, "bf16")
    got = m(x.to(DEV), ts.to(DEV), context=ctx.to(DEV))
    e = relerr(got, want)
    assert e <= EPS_TOL["bf16"], e


# ------------------------------------------------------------------------------------------
# the reference's other model configs (VERDICT r01 missing #3 / #4): config_train_stable.json (32-channel AutoencoderKL
# levels with 16 groups, head dim 32 in the UNet) and config_train_multigpu.json / config_optimized.json (attention at
# the 128-channel AutoencoderKL level + the decoder's non-local trio => d = 128 self-attention at (2s)^3 tokens)
# ------------------------------------------------------------------------------------------
STABLE_AE_KW = dict(spatial_dims=3, in_channels=1, out_channels=1, latent_channels=8, channels=[32, 64, 128],
                    num_res_blocks=1, norm_num_groups=16, norm_eps=1e-6, attention_levels=[False, False, True],
                    with_encoder_nonlocal_attn=False, with_decoder_nonlocal_attn=False)
STABLE_UNET_KW = dict(spatial_dims=3, in_channels=16, out_channels=8, channels=[128, 256, 512],
                      attention_levels=[False, False, True], num_head_channels=[0, 0, 32], num_res_blocks=1)
MULTIGPU_AE_KW = dict(spatial_dims=3, in_channels=1, out_channels=1, latent_channels=16, channels=[64, 128, 256],
                      num_res_blocks=2, norm_num_groups=32, norm_eps=1e-6, attention_levels=[False, True, True],
                      with_encoder_nonlocal_attn=False, with_decoder_nonlocal_attn=True)
OPTIMIZED_AE_KW = dict(MULTIGPU_AE_KW, channels=[128, 256, 256])


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name,kw,latent", [("stable", STABLE_AE_KW, (4, 6, 4)), ("multigpu", MULTIGPU_AE_KW, (6, 8, 4)),
                                            ("optimized", OPTIMIZED_AE_KW, (4, 4, 6))])
def test_autoencoder_other_reference_configs_vs_oracle(name, kw, latent, precision):
    import b200ldm
    ref = oracle.randomize_parameters(oracle.AutoencoderKL(**kw), 99).eval()
    m = port(ref, b200ldm.AutoencoderKL, kw, precision)
    z = torch.randn((2, kw["latent_channels"]) + latent, generator=gc.gen(31))
    img = torch.rand((1, 1) + tuple(4 * v for v in latent), generator=gc.gen(32))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        want_dec = ref.decode_stage_2_outputs(z)
        want_mu, want_sigma = ref.encode(img)
        want_rec = ref.reconstruct(img)
    tol = 1e-4 if precision == "fp32" else 2e-2
    e_dec = relerr(m.decode_stage_2_outputs(z.to(DEV)), want_dec)
    mu, sigma = m.encode(img.to(DEV))
    e_mu, e_rec = relerr(mu, want_mu), relerr(m.reconstruct(img.to(DEV)), want_rec)
    print(f"AE {name}/{precision}: decode {e_dec:.2e} mu {e_mu:.2e} reconstruct {e_rec:.2e}")
    assert e_dec <= tol and e_mu <= tol and e_rec <= 2 * tol, (e_dec, e_mu, e_rec)
    assert relerr(sigma, want_sigma) <= (1e-3 if precision == "fp32" else 1e-1)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("latent", [(8, 12, 8), (24, 24, 24)], ids=["12tokens", "216tokens"])
def test_unet_stable_config_head_dim_32_vs_oracle(latent, precision):
    """config_train_stable.json:40-49: channels (128, 256, 512), attention only at the last level with 32-channel heads
    (16 heads of head dim 32 - the tcgen05 kernel's 64-column tiles shared by head pairs), mid-block attention too"""
    import b200ldm
    ref = oracle.randomize_parameters(oracle.DiffusionModelUNet(**STABLE_UNET_KW), 77).eval()
    m = port(ref, b200ldm.DiffusionModelUNet, STABLE_UNET_KW, precision)
    x = torch.randn((2, 16) + latent, generator=gc.gen(33))
    ts = torch.tensor([700.0, 3.0])
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        want = ref(x, ts)
    e = relerr(m(x.to(DEV), ts.to(DEV)), want)
    print(f"unet stable {latent}/{precision}: eps relerr {e:.3e}")
    assert e <= EPS_TOL[precision], e


def test_stable_config_sampling_psnr():
    """config_train_stable end to end: DDPM-10 sampling with concat conditioning + decode, bf16 vs oracle >= 40 dB"""
    import b200ldm
    r_unet = oracle.randomize_parameters(oracle.DiffusionModelUNet(**STABLE_UNET_KW), 77).eval()
    r_ae = oracle.randomize_parameters(oracle.AutoencoderKL(**STABLE_AE_KW), 99).eval()
    unet = port(r_unet, b200ldm.DiffusionModelUNet, STABLE_UNET_KW, "bf16")
    ae = port(r_ae, b200ldm.AutoencoderKL, STABLE_AE_KW, "bf16")
    n, g = 10, gc.gen(34)
    shp = (2, 8, 8, 12, 8)
    x_T, cond = torch.randn(shp, generator=g), torch.randn(shp, generator=g)
    step_noise = torch.randn((n,) + shp, generator=g)
    s_ours, s_ref = gc.make_scheduler(b200ldm, "ddpm", n), gc.make_scheduler(oracle, "ddpm", n)
    img = b200ldm.LatentDiffusionInferer(s_ours, 1.0).sample(
        input_noise=x_T.to(DEV), autoencoder_model=ae, diffusion_model=unet, scheduler=s_ours,
        conditioning=cond.to(DEV), mode="concat", verbose=False, step_noise=step_noise.to(DEV))
    want = oracle.LatentDiffusionInferer(s_ref, 1.0).sample(x_T, r_ae, r_unet, s_ref, conditioning=cond, mode="concat",
                                                            verbose=False, step_noise=step_noise)
    p = psnr(img, want)
    print(f"config_train_stable DDPM-10 sampling + decode, bf16: {p:.1f} dB")
    assert img.shape == (2, 1, 32, 48, 32) and p >= 40.0, p
