"""Serving shim: the first consumer above the sampling path.

The reference's API server and Azure scoring script call a placeholder
(`DummyModel.generate(noise, num_inference_steps, guidance_scale)`, 3d_ldm/mlops/docker/api_server.py:138-155;
3d_ldm/mlops/deployment/score.py:116-126) once per requested sample with batch 1.  `LatentDiffusionGenerator` has
that surface - `.generate(noise, num_inference_steps=..., guidance_scale=...)` and `.forward(x)` returning a tensor
shaped like its argument - but runs the real path: DDPM / DDIM latent sampling with the b200ldm UNet + AutoencoderKL
decode.  Requests are batched: `generate_many` packs samples into fixed-size batches (the last one is padded) so every
batch replays the same captured CUDA graphs (LatentDiffusionInferer keeps them per shape / schedule).

Host code only; every arithmetic op is a libb200ldm kernel (no CPU fallback).
"""
from __future__ import annotations

import torch
from torch import nn

from .inferer import LatentDiffusionInferer
from .schedulers import DDIMScheduler, DDPMScheduler


def plan_batches(num_samples: int, max_batch: int) -> list[tuple[int, int]]:
    """[(first_sample, count)]: num_samples split into chunks of at most max_batch (each chunk runs PADDED to
    max_batch so the captured graph of that shape is reused; only `count` results are kept)."""
    if num_samples < 0 or max_batch <= 0:
        raise ValueError("num_samples must be >= 0 and max_batch > 0")
    return [(i, min(max_batch, num_samples - i)) for i in range(0, num_samples, max_batch)]


class LatentDiffusionGenerator(nn.Module):
    """generate(noise, num_inference_steps, guidance_scale) -> volumes shaped like `noise` ([B, 1, D, H, W]).

    autoencoder / unet: b200ldm modules on a CUDA device; scheduler_kwargs: as at 3d_ldm/inference.py:79-84;
    sampler: "ddpm" | "ddim"; conditioning_fn(batch, latent_spatial, device) -> [B, C_cond, d, h, w] for UNets trained
    with concat conditioning (None: zeros).  `guidance_scale` is accepted for interface parity and ignored - the
    reference's models are not trained with classifier-free guidance (its placeholder ignores it too).
    The content of `noise` is not used (it is image-shaped; the sampler needs latent-shaped noise): x_T is drawn from
    torch's global generator, which the reference's servers seed per request (`torch.manual_seed(request.seed)`).
    """

    def __init__(self, autoencoder, unet, scheduler_kwargs: dict, sampler: str = "ddpm", scale_factor: float = 1.0,
                 max_batch: int = 8, downsample: int = 4, conditioning_fn=None):
        super().__init__()
        if sampler not in ("ddpm", "ddim"):
            raise ValueError("sampler must be 'ddpm' or 'ddim'")
        self.autoencoder, self.unet = autoencoder, unet
        self.scheduler = (DDPMScheduler if sampler == "ddpm" else DDIMScheduler)(**scheduler_kwargs)
        self.inferer = LatentDiffusionInferer(self.scheduler, scale_factor=scale_factor)
        self.max_batch, self.downsample = max_batch, downsample
        self.conditioning_fn = conditioning_fn
        self.latent_channels = autoencoder.latent_channels
        self.cond_channels = unet.in_channels - self.latent_channels
        if self.cond_channels < 0:
            raise ValueError("UNet in_channels is smaller than the autoencoder's latent_channels")

    def _latent_spatial(self, spatial) -> tuple[int, ...]:
        if any(s % self.downsample for s in spatial):
            raise ValueError(f"volume extents {tuple(spatial)} must be multiples of {self.downsample}")
        return tuple(s // self.downsample for s in spatial)

    @torch.no_grad()
    def _run_batch(self, count: int, spatial, device, num_inference_steps: int) -> torch.Tensor:
        lat = self._latent_spatial(spatial)
        B = self.max_batch
        x_T = torch.randn((B, self.latent_channels) + lat, device=device)
        cond, mode = None, "crossattn"
        if self.cond_channels:
            mode = "concat"
            cond = (self.conditioning_fn(B, lat, device) if self.conditioning_fn is not None
                    else torch.zeros((B, self.cond_channels) + lat, device=device))
        if self.scheduler.num_inference_steps != num_inference_steps:
            self.scheduler.set_timesteps(num_inference_steps)
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())   # per-step noise stream, tied to the request's seed
        vols = self.inferer.sample(input_noise=x_T, autoencoder_model=self.autoencoder, diffusion_model=self.unet,
                                   scheduler=self.scheduler, conditioning=cond, mode=mode, verbose=False, seed=seed)
        return vols[:count]

    @torch.no_grad()
    def generate(self, noise: torch.Tensor, num_inference_steps: int = 100, guidance_scale: float = 7.5) -> torch.Tensor:
        if noise.ndim != 5:
            raise ValueError("noise must be [B, C, D, H, W]")
        if not noise.is_cuda:
            raise RuntimeError("b200ldm serving needs CUDA tensors (there is no CPU fallback)")
        outs = [self._run_batch(c, noise.shape[2:], noise.device, num_inference_steps)
                for _, c in plan_batches(noise.shape[0], self.max_batch)]
        out = torch.cat(outs) if outs else noise.new_zeros((0,) + tuple(noise.shape[1:]))
        return out.to(noise.dtype)

    def generate_many(self, num_samples: int, patch_size, device, num_inference_steps: int = 100,
                      seed: int | None = None) -> torch.Tensor:
        """What the request loop of api_server.py / score.py amounts to (`for i in range(num_samples)`), as batches."""
        if seed is not None:
            torch.manual_seed(seed)
        shape = (num_samples, self.autoencoder.out_channels) + tuple(patch_size)
        return self.generate(torch.empty(shape, device=device), num_inference_steps)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """`model(noise)` branch of the reference's request loop."""
        n = self.scheduler.num_inference_steps or self.scheduler.num_train_timesteps
        return self.generate(x, num_inference_steps=n)
