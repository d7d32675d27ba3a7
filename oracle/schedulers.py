"""Oracle schedulers: restatement of MONAI's DDPMScheduler / DDIMScheduler as used at
3d_ldm/inference.py:79-84.  TEST INFRASTRUCTURE ONLY.
All schedule tensors live on the CPU, exactly as upstream."""
from __future__ import annotations

import numpy as np
import torch


def _betas(schedule: str, T: int, beta_start: float = 1e-4, beta_end: float = 2e-2, **kw) -> torch.Tensor:
    if schedule == "linear_beta":
        return torch.linspace(beta_start, beta_end, T, dtype=torch.float32)
    if schedule == "scaled_linear_beta":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, T, dtype=torch.float32) ** 2
    if schedule == "sigmoid_beta":
        sig_range = kw.get("sig_range", 6)
        b = torch.linspace(-sig_range, sig_range, T)
        return torch.sigmoid(b) * (beta_end - beta_start) + beta_start
    raise ValueError(f"unknown schedule {schedule}")


class Scheduler:
    def __init__(self, num_train_timesteps: int = 1000, schedule: str = "linear_beta", **schedule_args):
        self.num_train_timesteps = num_train_timesteps
        self.betas = _betas(schedule, num_train_timesteps, **schedule_args)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())

    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        sa = ac[timesteps] ** 0.5
        sb = (1 - ac[timesteps]) ** 0.5
        while sa.ndim < original_samples.ndim:
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * original_samples + sb * noise


class DDPMScheduler(Scheduler):
    def __init__(self, num_train_timesteps=1000, schedule="linear_beta", variance_type="fixed_small",
                 clip_sample=True, prediction_type="epsilon", clip_sample_min=-1.0, clip_sample_max=1.0,
                 **schedule_args):
        super().__init__(num_train_timesteps, schedule, **schedule_args)
        if variance_type not in ("fixed_small", "fixed_large", "learned", "learned_range"):
            raise ValueError("Argument `variance_type` must be a member of `DDPMVarianceType`")
        if prediction_type not in ("epsilon", "sample", "v_prediction"):
            raise ValueError("Argument `prediction_type` must be a member of DDPMPredictionType")
        self.variance_type, self.clip_sample, self.prediction_type = variance_type, clip_sample, prediction_type
        self.clip_sample_values = [clip_sample_min, clip_sample_max]

    def set_timesteps(self, num_inference_steps, device=None):
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError("`num_inference_steps` cannot be larger than `num_train_timesteps`")
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts).to(device)

    def _get_variance(self, t, predicted_variance=None):
        """MONAI DDPMScheduler._get_variance [MONAI-recall]: the learned types use the second half of the model
        output; without one they fall through to the unclamped posterior variance."""
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[t - 1] if t > 0 else self.one
        var = (1 - a_prev) / (1 - a_t) * self.betas[t]
        if self.variance_type == "fixed_small":
            var = torch.clamp(var, min=1e-20)
        elif self.variance_type == "fixed_large":
            var = self.betas[t]
        elif self.variance_type == "learned" and predicted_variance is not None:
            return predicted_variance
        elif self.variance_type == "learned_range" and predicted_variance is not None:
            min_log, max_log = var, self.betas[t]
            frac = (predicted_variance + 1) / 2
            var = frac * max_log + (1 - frac) * min_log
        return var

    def step(self, model_output, timestep, sample, generator=None, noise=None):
        """`noise` (extension for parity tests): if given it replaces the
        internally drawn randn so both sides consume identical per-step noise."""
        t = int(timestep)
        predicted_variance = None
        if model_output.shape[1] == sample.shape[1] * 2 and self.variance_type in ("learned", "learned_range"):
            model_output, predicted_variance = torch.split(model_output, sample.shape[1], dim=1)
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[t - 1] if t > 0 else self.one  # t-1 even after set_timesteps(n<T)
        b_t, b_prev = 1 - a_t, 1 - a_prev
        if self.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        elif self.prediction_type == "sample":
            x0 = model_output
        else:
            x0 = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
        if self.clip_sample:
            x0 = torch.clamp(x0, self.clip_sample_values[0], self.clip_sample_values[1])
        c0 = (a_prev ** 0.5 * self.betas[t]) / b_t
        ct = self.alphas[t] ** 0.5 * b_prev / b_t
        prev = c0 * x0 + ct * sample
        if t > 0:
            if noise is None:
                noise = torch.randn(model_output.size(), dtype=model_output.dtype, layout=model_output.layout,
                                    generator=generator, device=model_output.device)
            prev = prev + (self._get_variance(t, predicted_variance) ** 0.5) * noise
        return prev, x0


class DDIMScheduler(Scheduler):
    def __init__(self, num_train_timesteps=1000, schedule="linear_beta", clip_sample=True,
                 set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon",
                 clip_sample_min=-1.0, clip_sample_max=1.0, **schedule_args):
        super().__init__(num_train_timesteps, schedule, **schedule_args)
        if prediction_type not in ("epsilon", "sample", "v_prediction"):
            raise ValueError("Argument `prediction_type` must be a member of DDIMPredictionType")
        self.prediction_type = prediction_type
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.clip_sample = clip_sample
        self.clip_sample_values = [clip_sample_min, clip_sample_max]
        self.steps_offset = steps_offset
        self.set_timesteps(self.num_train_timesteps)

    def set_timesteps(self, num_inference_steps, device=None):
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError("`num_inference_steps` cannot be larger than `num_train_timesteps`")
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        if self.steps_offset >= ratio:   # upstream raises here - also from the constructor's set_timesteps(T) for any offset > 0
            raise ValueError(f"`steps_offset`: {self.steps_offset} cannot be greater than or equal to "
                             f"`num_train_timesteps // num_inference_steps : {ratio}`")
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts).to(device)
        self.timesteps += self.steps_offset

    def _get_variance(self, t, prev_t):
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        return ((1 - a_prev) / (1 - a_t)) * (1 - a_t / a_prev)

    def step(self, model_output, timestep, sample, eta=0.0, generator=None, noise=None):
        t = int(timestep)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        if self.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
            eps = model_output
        elif self.prediction_type == "sample":
            x0 = model_output
            eps = (sample - a_t ** 0.5 * x0) / b_t ** 0.5
        else:
            x0 = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
            eps = (a_t ** 0.5) * model_output + (b_t ** 0.5) * sample
        if self.clip_sample:
            x0 = torch.clamp(x0, self.clip_sample_values[0], self.clip_sample_values[1])
        std = eta * self._get_variance(t, prev_t) ** 0.5
        prev = a_prev ** 0.5 * x0 + (1 - a_prev - std ** 2) ** 0.5 * eps
        if eta > 0:
            if noise is None:
                noise = torch.randn(model_output.shape, dtype=model_output.dtype, generator=generator,
                                    device=model_output.device)
            prev = prev + std * noise
        return prev, x0
