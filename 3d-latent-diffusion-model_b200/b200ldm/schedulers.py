"""B200 drop-ins for MONAI's DDPMScheduler / DDIMScheduler as constructed at
3d_ldm/inference.py:79-84.  The beta / alpha tables stay on the CPU like
upstream (a few hundred floats of host logic); `step()` runs the fused scheduler kernel
(b200ldm_scheduler_step) and `add_noise()` the add_noise kernel.  `coef_table()` exports the
per-step coefficient rows the graph-replayed sampling loop indexes by a device counter.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _betas(schedule: str, T: int, beta_start: float = 1e-4, beta_end: float = 2e-2, **kw) -> torch.Tensor:
    if schedule == "linear_beta":
        return torch.linspace(beta_start, beta_end, T, dtype=torch.float32)
    if schedule == "scaled_linear_beta":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, T, dtype=torch.float32) ** 2
    if schedule == "sigmoid_beta":
        sig_range = kw.get("sig_range", 6)
        return torch.sigmoid(torch.linspace(-sig_range, sig_range, T)) * (beta_end - beta_start) + beta_start
    raise ValueError(f"unknown noise schedule {schedule!r}")


def _randn_like_upstream(model_output, sample, generator) -> torch.Tensor:
    """upstream: torch.randn(model_output.size(), dtype=model_output.dtype, layout=..., generator=generator,
    device=model_output.device), taken after a learned-variance output has been split, i.e. always sample-shaped.
    Drawn on the generator's own device (upstream requires it to be the tensor's device), so the same generator state
    gives the same noise as MONAI."""
    z = torch.randn(tuple(sample.shape), dtype=model_output.dtype, generator=generator, device=generator.device)
    return z.to(sample.device)


class Scheduler:
    def __init__(self, num_train_timesteps: int = 1000, schedule: str = "linear_beta", **schedule_args):
        self.num_train_timesteps = num_train_timesteps
        self.betas = _betas(schedule, num_train_timesteps, **schedule_args)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())

    def _grid(self, num_inference_steps: int) -> np.ndarray:
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError(
                f"`num_inference_steps`: {num_inference_steps} cannot be larger than `self.num_train_timesteps`: "
                f"{self.num_train_timesteps}")
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        return (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)

    _PRED = ("epsilon", "sample", "v_prediction")

    def _x0_coefs(self, a_t):
        """(a_x, r): x0 = a_x * (x_t + r * model_output), or r * model_output when a_x == 0 (include/b200ldm.h)."""
        b_t = 1 - a_t
        if self.prediction_type == "epsilon":
            return 1.0 / a_t ** 0.5, -(b_t ** 0.5)
        if self.prediction_type == "sample":
            return torch.tensor(0.0), torch.tensor(1.0)
        return a_t ** 0.5, -(b_t ** 0.5) / a_t ** 0.5

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        t = timesteps.to("cpu").long().reshape(-1)
        n = original_samples.shape[0]
        if t.numel() == 1 and n > 1:
            t = t.expand(n)
        dev = original_samples.device
        sa = (self.alphas_cumprod[t] ** 0.5).to(dev, torch.float32).contiguous()
        sb = ((1 - self.alphas_cumprod[t]) ** 0.5).to(dev, torch.float32).contiguous()
        return ops.add_noise(original_samples.float().contiguous(), noise.float().contiguous(), sa, sb)

    # ---- fused-step plumbing
    def coef_row(self, timestep: int) -> list[float]:
        raise NotImplementedError

    def _state_key(self) -> tuple:
        """everything a coefficient row depends on besides the timestep (attributes may be reassigned after construction)"""
        return (self.num_inference_steps, self.prediction_type, bool(self.clip_sample), tuple(self.clip_sample_values),
                getattr(self, "variance_type", None), getattr(self, "eta", None), getattr(self, "steps_offset", None))

    def _memo_table(self, name: str, row_fn, device) -> torch.Tensor:
        """[n_steps, k] table on `device`, built once per (scheduler state, timestep grid): a cached sample() call used to
        redo ~1000 tensor-scalar operations here (each a device sync when `timesteps` lives on the GPU)"""
        ts = tuple(torch.as_tensor(self.timesteps).tolist())
        key = (name, str(device), ts, self._state_key())
        memo = self.__dict__.setdefault("_table_memo", {})
        t = memo.get(key)
        if t is None:
            if len(memo) >= 8:
                memo.clear()
            t = torch.tensor([row_fn(int(v)) for v in ts], dtype=torch.float32).to(device).contiguous()
            memo[key] = t
        return t

    def coef_table(self, device) -> torch.Tensor:
        """[n_steps, 8] coefficient rows; memoised - callers copy it into their own buffers, never write it"""
        return self._memo_table("coef", self.coef_row, device)

    learned_variance = False   # DDPMScheduler(variance_type="learned" | "learned_range") sets it

    def var_coef_row(self, timestep: int) -> list[float]:
        raise NotImplementedError

    def var_coef_table(self, device) -> torch.Tensor:
        """[n_steps, 2] rows {A, B}: per-element noise scale sqrt(A * predicted_variance + B) (include/b200ldm.h)."""
        return self._memo_table("var_coef", self.var_coef_row, device)

    def _step_kernel(self, model_output, timestep, sample, noise):
        """NCDHW fp32 tensors in, (prev_sample, pred_original_sample) out."""
        dev = sample.device
        N, Cc = sample.shape[:2]
        cpad = ops.round_up(Cc, 4)
        split = self.learned_variance and model_output.shape[1] == 2 * Cc \
            and tuple(model_output.shape[2:]) == tuple(sample.shape[2:])
        if not split and model_output.shape != sample.shape:
            raise ValueError("model_output and sample must have the same shape (twice the channels only with a learned "
                             "variance_type)")
        coef = self._coef_on(dev, int(timestep), split=split)
        x = ops.to_ndhwc(sample, cpad, torch.float32)
        var_coef = None
        if split:   # [eps | predicted variance] side by side, each half padded to cpad channels
            e = ops.to_ndhwc(model_output[:, :Cc], 2 * cpad, torch.float32)
            ops.to_ndhwc(model_output[:, Cc:], 2 * cpad, torch.float32, out=e, c_offset=cpad)
            var_coef = torch.tensor([self.var_coef_row(int(timestep))], dtype=torch.float32).to(dev)
        else:
            e = ops.to_ndhwc(model_output, cpad, torch.float32)
        nz = None if noise is None else ops.to_ndhwc(noise, cpad, torch.float32)
        x0 = torch.empty_like(x)
        ops.scheduler_step(e, x, coef, None, cpad, noise=nz, x0_out=x0, seed=self._seed(), var_offset=cpad if split else 0,
                           var_coef=var_coef)
        return ops.to_ncdhw(x, Cc), ops.to_ncdhw(x0, Cc)

    def _coef_on(self, dev, timestep: int, eta=None, split: bool = False) -> torch.Tensor:
        """coefficient row of one timestep on the device, cached: a fresh pageable upload per step() would block the
        host until the stream drains"""
        key = (str(dev), timestep, eta, split, self._state_key())
        cache = self.__dict__.setdefault("_coef_cache", {})
        t = cache.get(key)
        if t is None:
            if len(cache) > 4096:
                cache.clear()
            t = torch.tensor([self.coef_row(timestep)], dtype=torch.float32).to(dev)
            cache[key] = t
        return t

    _seed_state = 0

    def _seed(self) -> int:
        Scheduler._seed_state += 1
        return (int(torch.initial_seed()) * 0x9E3779B97F4A7C15 + Scheduler._seed_state) & (2 ** 64 - 1)


class DDPMScheduler(Scheduler):
    def __init__(self, num_train_timesteps: int = 1000, schedule: str = "linear_beta", variance_type: str = "fixed_small",
                 clip_sample: bool = True, prediction_type: str = "epsilon", clip_sample_min: float = -1.0,
                 clip_sample_max: float = 1.0, **schedule_args):
        super().__init__(num_train_timesteps, schedule, **schedule_args)
        if variance_type not in ("fixed_small", "fixed_large", "learned", "learned_range"):
            raise ValueError("Argument `variance_type` must be a member of `DDPMVarianceType`")
        if prediction_type not in self._PRED:
            raise ValueError("Argument `prediction_type` must be a member of DDPMPredictionType")
        self.learned_variance = variance_type in ("learned", "learned_range")
        self.variance_type, self.clip_sample, self.prediction_type = variance_type, clip_sample, prediction_type
        self.clip_sample_values = [clip_sample_min, clip_sample_max]

    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        self.timesteps = torch.from_numpy(self._grid(num_inference_steps)).to(device)

    def _get_variance(self, t: int) -> torch.Tensor:
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[t - 1] if t > 0 else self.one
        var = (1 - a_prev) / (1 - a_t) * self.betas[t]
        if self.variance_type == "fixed_small":
            return torch.clamp(var, min=1e-20)
        # learned types WITHOUT a predicted variance fall through to the unclamped posterior variance (upstream)
        return self.betas[t] if self.variance_type == "fixed_large" else var

    def var_coef_row(self, t: int) -> list[float]:
        """{A, B}: the model's predicted variance v gives the noise scale sqrt(A*v + B).  learned: the prediction IS the
        variance; learned_range: v in [-1, 1] interpolates var_t .. beta_t, frac = (v + 1) / 2."""
        if t == 0:
            return [0.0, 0.0]
        if self.variance_type == "learned":
            return [1.0, 0.0]
        a_t, a_prev = self.alphas_cumprod[t], self.alphas_cumprod[t - 1]
        lo, hi = (1 - a_prev) / (1 - a_t) * self.betas[t], self.betas[t]
        return [float((hi - lo) / 2), float((hi + lo) / 2)]

    def coef_row(self, t: int) -> list[float]:
        """{a_x, r, c_x0, c_xt, c_e, sigma, clip_lo, clip_hi} (include/b200ldm.h) in fp32 torch arithmetic that mirrors
        upstream's operation order (alphas_cumprod[t-1] even after set_timesteps(n<T))."""
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[t - 1] if t > 0 else self.one
        b_t, b_prev = 1 - a_t, 1 - a_prev
        c0 = (a_prev ** 0.5 * self.betas[t]) / b_t
        ct = self.alphas[t] ** 0.5 * b_prev / b_t
        sigma = self._get_variance(t) ** 0.5 if t > 0 else torch.tensor(0.0)
        lo, hi = (self.clip_sample_values if self.clip_sample else (-3.0e38, 3.0e38))
        a_x, a_e = self._x0_coefs(a_t)
        return [float(a_x), float(a_e), float(c0), float(ct), 0.0, float(sigma), float(lo), float(hi)]

    def step(self, model_output, timestep, sample, generator=None, noise=None):
        """Same contract as upstream.  `generator`: this step's N(0,I) is drawn from it with torch.randn exactly as
        upstream does (same generator state -> same noise as MONAI).  `noise` (extension) supplies the draw itself
        (parity tests).  With neither, the kernel draws Philox noise seeded from torch's global seed."""
        t = int(timestep)
        if noise is None and generator is not None and t > 0:
            noise = _randn_like_upstream(model_output, sample, generator)
        return self._step_kernel(model_output, t, sample, noise if t > 0 else None)


class DDIMScheduler(Scheduler):
    def __init__(self, num_train_timesteps: int = 1000, schedule: str = "linear_beta", clip_sample: bool = True,
                 set_alpha_to_one: bool = True, steps_offset: int = 0, prediction_type: str = "epsilon",
                 clip_sample_min: float = -1.0, clip_sample_max: float = 1.0, **schedule_args):
        super().__init__(num_train_timesteps, schedule, **schedule_args)
        if prediction_type not in self._PRED:
            raise ValueError("Argument `prediction_type` must be a member of DDIMPredictionType")
        self.prediction_type = prediction_type
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.clip_sample = clip_sample
        self.clip_sample_values = [clip_sample_min, clip_sample_max]
        self.steps_offset = steps_offset
        self.eta = 0.0
        self.set_timesteps(self.num_train_timesteps)

    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        grid = self._grid(num_inference_steps)
        ratio = self.num_train_timesteps // num_inference_steps
        if self.steps_offset >= ratio:
            raise ValueError(
                f"`steps_offset`: {self.steps_offset} cannot be greater than or equal to "
                f"`num_train_timesteps // num_inference_steps : {ratio}` as this will cause timesteps to exceed"
                f" the max train timestep.")
        self.timesteps = torch.from_numpy(grid).to(device)
        self.timesteps += self.steps_offset

    def _get_variance(self, t: int, prev_t: int) -> torch.Tensor:
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        return ((1 - a_prev) / (1 - a_t)) * (1 - a_t / a_prev)

    def coef_row(self, t: int, eta: float | None = None) -> list[float]:
        eta = self.eta if eta is None else eta
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        std = eta * self._get_variance(t, prev_t) ** 0.5
        c_dir = (1 - a_prev - std ** 2) ** 0.5        # x' = sqrt(a_prev)*x0 + c_dir*eps_hat + std*z
        # eps_hat (from the UNCLIPPED x0, as upstream) is linear in (x_t, model output m):
        #   epsilon: m;  sample: (x_t - sqrt(a_t) m)/sqrt(b_t);  v_prediction: sqrt(a_t) m + sqrt(b_t) x_t
        if self.prediction_type == "epsilon":
            c_xt, c_e = torch.tensor(0.0), c_dir
        elif self.prediction_type == "sample":
            c_xt, c_e = c_dir / b_t ** 0.5, -c_dir * a_t ** 0.5 / b_t ** 0.5
        else:
            c_xt, c_e = c_dir * b_t ** 0.5, c_dir * a_t ** 0.5
        lo, hi = (self.clip_sample_values if self.clip_sample else (-3.0e38, 3.0e38))
        a_x, a_e = self._x0_coefs(a_t)
        return [float(a_x), float(a_e), float(a_prev ** 0.5), float(c_xt), float(c_e), float(std), float(lo), float(hi)]

    def step(self, model_output, timestep, sample, eta: float = 0.0, generator=None, noise=None):
        t = int(timestep)
        if noise is None and generator is not None and eta > 0:
            noise = _randn_like_upstream(model_output, sample, generator)
        old = self.eta
        self.eta = eta
        try:
            return self._step_kernel(model_output, t, sample, noise if eta > 0 else None)
        finally:
            self.eta = old
