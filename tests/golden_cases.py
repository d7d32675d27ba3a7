"""Seeded tiny cases shared by the golden-vector generator, the oracle tests and the GPU parity tests.
Shapes follow the reference's smallest config family (channels multiples of 64; SURVEY.md 8c pin 3)."""
import torch

import oracle

UNET_KW = dict(spatial_dims=3, in_channels=8, out_channels=4, channels=(64, 128), attention_levels=(False, True),
               num_head_channels=(0, 64), num_res_blocks=1, norm_num_groups=32, norm_eps=1e-6)
AE_KW = dict(spatial_dims=3, in_channels=1, out_channels=1, latent_channels=4, channels=(64, 128),
             num_res_blocks=1, norm_num_groups=32, norm_eps=1e-6, attention_levels=(False, True),
             with_encoder_nonlocal_attn=False, with_decoder_nonlocal_attn=False)
LATENT = (6, 8, 4)  # deliberately non-cubic; image = 2x (one AE down-sampling level)
# cross-attention conditioning (with_conditioning=True -> SpatialTransformer at the attention level),
# two transformer layers, a context width that is NOT a multiple of 64, class embeddings on top
XATTN_UNET_KW = dict(spatial_dims=3, in_channels=4, out_channels=4, channels=(64, 128), attention_levels=(False, True),
                     num_head_channels=(0, 64), num_res_blocks=1, norm_num_groups=32, norm_eps=1e-6,
                     with_conditioning=True, cross_attention_dim=24, transformer_num_layers=2, num_class_embeds=5)
XATTN_CTX_LEN = 7


def gen(seed):
    return torch.Generator(device="cpu").manual_seed(seed)


def tiny_unet(seed=1234):
    return oracle.randomize_parameters(oracle.DiffusionModelUNet(**UNET_KW), seed).eval()


def tiny_ae(seed=4321):
    return oracle.randomize_parameters(oracle.AutoencoderKL(**AE_KW), seed).eval()


def tiny_unet_case():
    x = torch.randn((2, 8) + LATENT, generator=gen(7))
    ts = torch.tensor([37.0, 912.0])
    return tiny_unet(), x, ts


def tiny_xattn_unet(seed=2468, **over):
    return oracle.randomize_parameters(oracle.DiffusionModelUNet(**{**XATTN_UNET_KW, **over}), seed).eval()


def tiny_xattn_case():
    """(oracle UNet, x, timesteps, context [2, 7, 24], class labels)"""
    g = gen(11)
    x = torch.randn((2, 4) + LATENT, generator=g)
    ctx = torch.randn((2, XATTN_CTX_LEN, 24), generator=g)
    return tiny_xattn_unet(), x, torch.tensor([11.0, 640.0]), ctx, torch.tensor([3, 0])


def tiny_xattn_sampling_case(kind="ddpm", n_steps=4):
    """mode="crossattn" sampling with a context (no class embeddings: MONAI's sample() cannot pass labels)."""
    unet, ae = tiny_xattn_unet(num_class_embeds=None), tiny_ae()
    sched = make_scheduler(oracle, kind, n_steps)
    x_T, _, step_noise = sampling_inputs(n_steps)
    ctx = torch.randn((2, XATTN_CTX_LEN, 24), generator=gen(12))
    inf = oracle.LatentDiffusionInferer(sched, scale_factor=1.0)
    image, latent = inf.sample(x_T, ae, unet, sched, conditioning=ctx, mode="crossattn", verbose=False,
                               step_noise=step_noise, return_latent=True)
    return (unet, ae, x_T, ctx, step_noise), image.contiguous(), latent.contiguous()


def tiny_ae_case():
    return tiny_ae(), torch.randn((2, 4) + LATENT, generator=gen(8))


def tiny_ae_image():
    return torch.rand((1, 1) + tuple(2 * v for v in LATENT), generator=gen(9))


def sampling_inputs(n_steps, batch=2):
    g = gen(42)
    x_T = torch.randn((batch, 4) + LATENT, generator=g)
    cond = torch.randn((batch, 4) + LATENT, generator=g)
    step_noise = torch.randn((n_steps, batch, 4) + LATENT, generator=g)
    return x_T, cond, step_noise


def make_scheduler(mod, kind, n_steps, **extra):
    """mod: `oracle` or `b200ldm` - same constructor surface (3d_ldm/inference.py:79-84)."""
    cls = mod.DDPMScheduler if kind == "ddpm" else mod.DDIMScheduler
    s = cls(num_train_timesteps=1000, schedule="scaled_linear_beta", beta_start=0.0015, beta_end=0.0195, **extra)
    s.set_timesteps(n_steps)
    return s


def tiny_sampling_case(kind, n_steps=5):
    unet, ae = tiny_unet(), tiny_ae()
    sched = make_scheduler(oracle, kind, n_steps)
    x_T, cond, step_noise = sampling_inputs(n_steps)
    inf = oracle.LatentDiffusionInferer(sched, scale_factor=1.0)
    image, latent = inf.sample(x_T, ae, unet, sched, conditioning=cond, mode="concat", verbose=False,
                               step_noise=step_noise, return_latent=True)
    return image.contiguous(), latent.contiguous()
