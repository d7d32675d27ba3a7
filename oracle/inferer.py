"""Oracle inferers: restatement of MONAI DiffusionInferer / LatentDiffusionInferer as
called at 3d_ldm/inference.py:85,94-99 and 3d_ldm/train_diffusion.py:197-205,326-333
(SURVEY.md Appendix D).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import torch


class DiffusionInferer:
    def __init__(self, scheduler):
        self.scheduler = scheduler

    @torch.no_grad()
    def sample(self, input_noise, diffusion_model, scheduler=None, save_intermediates=False,
               intermediate_steps=100, conditioning=None, mode="crossattn", verbose=True, seg=None,
               step_noise=None):
        """`step_noise` (parity-test extension): tensor [n_steps, *input_noise.shape];
        row i is consumed by the i-th scheduler.step instead of a fresh randn."""
        if mode not in ("crossattn", "concat"):
            raise NotImplementedError(f"{mode} condition is not supported")
        if mode == "concat" and conditioning is None:
            raise ValueError("Conditioning must be supplied for if condition mode is concat.")
        scheduler = scheduler or self.scheduler
        image = input_noise
        intermediates = []
        for i, t in enumerate(scheduler.timesteps):
            ts = torch.Tensor((t,)).to(input_noise.device)
            if mode == "concat":
                eps = diffusion_model(torch.cat([image, conditioning], dim=1), timesteps=ts, context=None)
            else:
                eps = diffusion_model(image, timesteps=ts, context=conditioning)
            if step_noise is not None:
                image, _ = scheduler.step(eps, t, image, noise=step_noise[i])
            else:
                image, _ = scheduler.step(eps, t, image)
            if save_intermediates and t % intermediate_steps == 0:
                intermediates.append(image)
        return (image, intermediates) if save_intermediates else image

    def __call__(self, inputs, diffusion_model, noise, timesteps, condition=None, mode="crossattn", seg=None):
        if mode not in ("crossattn", "concat"):
            raise NotImplementedError(f"{mode} condition is not supported")
        noisy = self.scheduler.add_noise(original_samples=inputs, noise=noise, timesteps=timesteps)
        if mode == "concat":
            if condition is None:
                raise ValueError("Conditioning is required for concat condition")
            noisy = torch.cat([noisy, condition], dim=1)
            condition = None
        return diffusion_model(x=noisy, timesteps=timesteps, context=condition)


class LatentDiffusionInferer(DiffusionInferer):
    def __init__(self, scheduler, scale_factor: float = 1.0, ldm_latent_shape=None, autoencoder_latent_shape=None):
        super().__init__(scheduler)
        self.scale_factor = scale_factor

    @torch.no_grad()
    def sample(self, input_noise, autoencoder_model, diffusion_model, scheduler=None, save_intermediates=False,
               intermediate_steps=100, conditioning=None, mode="crossattn", verbose=True, seg=None,
               step_noise=None, return_latent=False):
        out = super().sample(input_noise, diffusion_model, scheduler, save_intermediates, intermediate_steps,
                             conditioning, mode, verbose, seg, step_noise=step_noise)
        latent, lat_inter = out if save_intermediates else (out, None)
        image = autoencoder_model.decode_stage_2_outputs(latent / self.scale_factor)
        if save_intermediates:
            inter = [autoencoder_model.decode_stage_2_outputs(l / self.scale_factor) for l in lat_inter]
            return image, inter
        if return_latent:
            return image, latent
        return image

    def __call__(self, inputs, autoencoder_model, diffusion_model, noise, timesteps, condition=None,
                 mode="crossattn", seg=None):
        with torch.no_grad():
            latent = autoencoder_model.encode_stage_2_inputs(inputs) * self.scale_factor
        return super().__call__(latent, diffusion_model, noise, timesteps, condition, mode, seg)
