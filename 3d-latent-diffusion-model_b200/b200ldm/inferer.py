"""B200 drop-ins for MONAI's DiffusionInferer / LatentDiffusionInferer:

    inferer = LatentDiffusionInferer(scheduler, scale_factor=1.0)            3d_ldm/inference.py:85
    inferer.sample(input_noise=..., autoencoder_model=..., diffusion_model=..., scheduler=...)   :94-99
    inferer.sample(..., conditioning=latents, mode="concat")                 3d_ldm/train_diffusion.py:326-333
    inferer(inputs=..., noise=..., timesteps=..., condition=..., mode="concat")  3d_ldm/train_diffusion.py:197-205

`sample` is the hot path.  When the models are b200ldm modules it runs the fused loop: the latent
stays in NDHWC on the device (fp32 master + a compute-dtype copy inside the UNet input buffer),
time-embedding projections for every step are tabulated once, and ONE CUDA graph
(select temb row -> UNet -> fused scheduler update -> counter++) is replayed per step; per-step
coefficients / temb rows / noise slabs are indexed by a device-side step counter.
"""
from __future__ import annotations

import os

import torch

from . import ops
from .nets import AutoencoderKL, DiffusionModelUNet, _PREC, _pad64, param_signature

USE_GRAPH = os.environ.get("B200LDM_CUDA_GRAPH", "1") != "0"
FUSE_SCHED = os.environ.get("B200LDM_FUSE_SCHED", "1") != "0"
# The volumes of a batch are independent, so the TOP resolution level of the UNet (conv_in, down block 0, the last up
# block, the output conv) runs as BATCH_STREAMS sub-batches on parallel streams = parallel branches of the captured
# graph (DiffusionModelUNet.forward_ndhwc(split=...)): one sub-batch's CTAs fill the SMs another leaves idle in the
# partial last wave of a conv (8 x 20^3: 256 CTAs on 148 SMs = 2 waves for 1.7 waves of work).  Splitting the WHOLE
# step was measured slower (profiles/r01d_batch_streams_ab.log: 6.17 -> 6.28 ms with 2 streams - the small levels need
# more split-K and re-read the weights), so the lower levels stay on the whole batch.  The noise does not depend on
# the split.
BATCH_STREAMS = int(os.environ.get("B200LDM_BATCH_STREAMS", "2"))


class SamplingLoop:
    """Device-resident state + (optionally graph-captured) body of one denoising step.

    All state lives in buffers allocated once per (model, shapes, timesteps); `reset` refills them in place, so the
    captured graph (whose kernels hold these pointers) serves every later batch of the same shape - including a new
    seed, which the scheduler kernel reads from device memory."""

    def __init__(self, unet: DiffusionModelUNet, scheduler, input_noise: torch.Tensor, conditioning, mode: str,
                 step_noise=None, seed: int = 0, use_graph: bool = True, batch_streams: int | None = None):
        dev = input_noise.device
        self.unet = unet
        self.dtype = _PREC[unet.precision]
        N, self.c_lat = input_noise.shape[:2]
        self.shape = tuple(input_noise.shape)
        self.spatial = tuple(input_noise.shape[2:])
        self.mode = mode
        self.c_in = self.c_lat + (conditioning.shape[1] if mode == "concat" else 0)
        # mode "crossattn": the conditioning is the cross-attention context [N, Lc, cross_attention_dim]; its padded
        # compute-dtype token matrix lives in a fixed buffer the captured step reads
        self.ctx = None
        if mode == "crossattn" and conditioning is not None:
            if not unet.with_conditioning:
                raise ValueError("model should have with_conditioning = True if context is provided")
            self.ctx = unet.context_tokens(conditioning.to(dev))
        if self.c_in != unet.in_channels:
            raise ValueError(f"UNet expects {unet.in_channels} input channels, sampling provides {self.c_in}")
        if self.c_lat % 4:
            raise ValueError("latent channel count must be a multiple of 4")
        self.n_steps = len(scheduler.timesteps)
        # state
        self.x_t = ops.to_ndhwc(input_noise, self.c_lat, torch.float32)            # fp32 master [N,D,H,W,C]
        self.x_model = ops.to_ndhwc(input_noise, _pad64(self.c_in), self.dtype)     # UNet input buffer (padding = 0)
        ts = torch.as_tensor(scheduler.timesteps).to(device=dev, dtype=torch.float32)
        self.temb_table = unet.time_embedding_table(ts)                            # [n_steps, total]
        self.temb_cur = torch.empty((1, self.temb_table.shape[1]), dtype=torch.float32, device=dev)
        self.coef = scheduler.coef_table(dev).clone()
        # learned variance (DDPMScheduler variance_type "learned" / "learned_range"): the UNet predicts 2 * c_lat channels
        # and the second half scales each element's noise; that update runs in the standalone scheduler kernel
        self.var_coef = None
        if unet.out_channels == 2 * self.c_lat and getattr(scheduler, "learned_variance", False):
            self.var_coef = scheduler.var_coef_table(dev).clone()
        elif unet.out_channels != self.c_lat:
            raise ValueError(f"UNet predicts {unet.out_channels} channels for a {self.c_lat}-channel latent "
                             "(twice the channels needs a learned variance_type)")
        self.counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.noise, self.noise_stride = None, 0
        if step_noise is not None:
            self.noise = torch.empty((self.n_steps * N,) + self.spatial + (self.c_lat,), dtype=torch.float32, device=dev)
            self.noise_stride = self.noise.numel() // self.n_steps
        self.graph = None
        self.launches_per_step = 0
        self.use_graph = use_graph and USE_GRAPH
        # sub-batches [lo, hi) of the batch axis, one stream each (only worth it inside a captured graph and where the
        # top level is big enough to fill the chip more than once: >= 2 x 148 tiles of 128 voxels)
        n_sub = BATCH_STREAMS if batch_streams is None else batch_streams
        if batch_streams is None and input_noise.numel() // self.c_lat < 2 * 148 * 128:
            n_sub = 1
        n_sub = max(1, min(n_sub, N)) if self.use_graph else 1
        cuts = [round(i * N / n_sub) for i in range(n_sub + 1)]
        self.subs = [(cuts[i], cuts[i + 1]) for i in range(n_sub)]
        self.side_streams = [torch.cuda.Stream(device=dev) for _ in range(n_sub - 1)]
        self._sched = [None] * n_sub
        self.reset(input_noise, conditioning, step_noise, seed, scheduler, _fresh=True)

    def reset(self, input_noise, conditioning, step_noise, seed: int, scheduler, _fresh: bool = False):
        """Start a new trajectory in the existing buffers (same shapes / timesteps as at construction)."""
        if tuple(input_noise.shape) != self.shape or len(scheduler.timesteps) != self.n_steps:
            raise ValueError("SamplingLoop.reset: shapes / number of steps differ from the captured loop")
        if (step_noise is not None) != (self.noise is not None):
            raise ValueError("SamplingLoop.reset: step_noise presence differs from the captured loop")
        if not _fresh:
            ops.to_ndhwc(input_noise, self.c_lat, torch.float32, out=self.x_t)
            ops.to_ndhwc(input_noise, _pad64(self.c_in), self.dtype, out=self.x_model)
            self.coef.copy_(scheduler.coef_table(self.coef.device))
            if self.var_coef is not None:
                self.var_coef.copy_(scheduler.var_coef_table(self.coef.device))
        if self.mode == "concat":
            ops.to_ndhwc(conditioning, _pad64(self.c_in), self.dtype, out=self.x_model, c_offset=self.c_lat)
        elif self.ctx is not None and not _fresh:
            self.unet.context_tokens(conditioning.to(self.ctx.device), out=self.ctx)
        if step_noise is not None:
            if step_noise.shape[0] != self.n_steps or tuple(step_noise.shape[1:]) != self.shape:
                raise ValueError("step_noise must be [n_steps, *input_noise.shape]")
            sn = step_noise.to(self.x_t.device).reshape((self.n_steps * self.shape[0],) + self.shape[1:])
            ops.to_ndhwc(sn, self.c_lat, torch.float32, out=self.noise)
        self.seed_dev.fill_(int(seed) & (2 ** 63 - 1))
        self.counter.zero_()

    def _sub_sched(self, i: int):
        """scheduler parameters of sub-batch i: batch slices of the state buffers (contiguous views), the Philox index
        of its first element so the noise does not depend on the split"""
        if self._sched[i] is None:
            lo, hi = self.subs[i]
            noise = None if self.noise is None else self.noise[lo:]      # slab of step s = rows [s * N + lo, ...)
            self._sched[i] = ops.make_sched_params(self.x_t[lo:hi], self.coef, self.counter, self.c_lat, self.dtype,
                                                   noise=noise, noise_step_stride=self.noise_stride,
                                                   x_model=self.x_model[lo:hi], seed=self.seed_dev,
                                                   philox_offset=lo * (self.x_t[0].numel() // 4))
        return self._sched[i]

    def _body(self):
        ops.select_row(self.temb_table, self.counter, self.temb_cur)
        if FUSE_SCHED and self.var_coef is None and self.unet.out[2].conv.tap_expanded(self.x_model):
            # the scheduler update runs inside the UNet's last kernel (the output conv's tap gather) on the fp32 sums
            if len(self.subs) > 1:
                self.unet.forward_ndhwc(self.x_model, self.temb_cur, 0, ctx=self.ctx,
                                        sched=[self._sub_sched(i) for i in range(len(self.subs))],
                                        split=(self.subs, self.side_streams))
            else:
                self.unet.forward_ndhwc(self.x_model, self.temb_cur, 0, sched=self._sub_sched(0), ctx=self.ctx)
        else:
            eps = self.unet.forward_ndhwc(self.x_model, self.temb_cur, 0, ctx=self.ctx)
            ops.scheduler_step(eps, self.x_t, self.coef, self.counter, self.c_lat, noise=self.noise,
                               noise_step_stride=self.noise_stride, x_model=self.x_model, seed=self.seed_dev,
                               var_offset=self.c_lat if self.var_coef is not None else 0, var_coef=self.var_coef)
        ops.counter_add(self.counter, 1)

    def _capture(self):
        # warm-up (weight repack, cudaFuncSetAttribute, allocator) on a scratch copy of the state
        keep_x, keep_m = self.x_t.clone(), self.x_model.clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._body()
        torch.cuda.current_stream().wait_stream(s)
        self.x_t.copy_(keep_x)
        self.x_model.copy_(keep_m)
        self.counter.zero_()
        g = torch.cuda.CUDAGraph()
        l0 = ops.launch_count()
        with torch.cuda.graph(g):
            self._body()
        self.launches_per_step = ops.launch_count() - l0
        self.x_t.copy_(keep_x)
        self.x_model.copy_(keep_m)
        self.counter.zero_()
        self.graph = g

    def step(self):
        if self.use_graph and USE_GRAPH:
            if self.graph is None:
                self._capture()
            self.graph.replay()
        else:
            self._body()

    def latent_ncdhw(self) -> torch.Tensor:
        return ops.to_ncdhw(self.x_t, self.c_lat)


class DiffusionInferer:
    MAX_CACHED_LOOPS = 2   # each holds a CUDA graph + its private activation pool

    def __init__(self, scheduler):
        self.scheduler = scheduler
        self._loops = {}

    def _loop_for(self, unet, scheduler, input_noise, conditioning, mode, step_noise, seed):
        """Cached SamplingLoop (buffers + captured graph) for this model / shape / schedule; a repeated `sample` call
        then costs T graph replays and nothing else (capture and its allocator traffic were 0.2-1.1 s per call)."""
        key = (id(unet), param_signature(unet), unet.precision, tuple(input_noise.shape), str(input_noise.device),
               None if conditioning is None else tuple(conditioning.shape), mode,
               tuple(torch.as_tensor(scheduler.timesteps).tolist()), step_noise is not None,   # one transfer, not T
               getattr(scheduler, "variance_type", None))
        loop = self._loops.get(key)
        if loop is None:
            while len(self._loops) >= self.MAX_CACHED_LOOPS:
                self._loops.pop(next(iter(self._loops)))
            loop = SamplingLoop(unet, scheduler, input_noise, conditioning, mode, step_noise=step_noise, seed=seed,
                                use_graph=len(scheduler.timesteps) > 2)
            self._loops[key] = loop
        else:
            loop.reset(input_noise, conditioning, step_noise, seed, scheduler)
        return loop

    @torch.no_grad()
    def sample(self, input_noise, diffusion_model, scheduler=None, save_intermediates=False, intermediate_steps=100,
               conditioning=None, mode="crossattn", verbose=True, seg=None, step_noise=None, seed=None,
               _return_loop=False):
        if mode not in ("crossattn", "concat"):
            raise NotImplementedError(f"{mode} condition is not supported")
        if mode == "concat" and conditioning is None:
            raise ValueError("Conditioning must be supplied for if condition mode is concat.")
        scheduler = scheduler or self.scheduler
        if not isinstance(diffusion_model, DiffusionModelUNet):
            raise TypeError("b200ldm inferers drive b200ldm.DiffusionModelUNet (there is no generic / CPU fallback)")
        if mode == "crossattn" and conditioning is not None and not diffusion_model.with_conditioning:
            raise ValueError("model should have with_conditioning = True if context is provided")
        if not input_noise.is_cuda:
            raise RuntimeError("b200ldm sampling needs CUDA tensors (there is no CPU fallback)")
        if seed is None:
            # a fresh noise stream per call, drawn from (and advancing) torch's global generator like the reference's
            # per-step torch.randn does: torch.manual_seed(s) still makes a whole run reproducible, and two consecutive
            # calls never share their z_1..z_T (torch.initial_seed() is constant for the process - ADVICE r01)
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        loop = self._loop_for(diffusion_model, scheduler, input_noise, conditioning, mode, step_noise, seed)
        intermediates = []
        for t in torch.as_tensor(scheduler.timesteps).tolist():
            loop.step()
            if save_intermediates and t % intermediate_steps == 0:
                intermediates.append(loop.latent_ncdhw())
        image = loop.latent_ncdhw()
        if _return_loop:
            return image, intermediates, loop
        return (image, intermediates) if save_intermediates else image

    @torch.no_grad()
    def __call__(self, inputs, diffusion_model, noise, timesteps, condition=None, mode="crossattn", seg=None):
        if mode not in ("crossattn", "concat"):
            raise NotImplementedError(f"{mode} condition is not supported")
        noisy = self.scheduler.add_noise(original_samples=inputs, noise=noise, timesteps=timesteps)
        if mode == "concat":
            if condition is None:
                raise ValueError("Conditioning is required for concat condition")
            noisy = torch.cat([noisy, condition.to(noisy.dtype)], dim=1)
            condition = None
        return diffusion_model(x=noisy, timesteps=timesteps, context=condition)


class LatentDiffusionInferer(DiffusionInferer):
    def __init__(self, scheduler, scale_factor: float = 1.0, ldm_latent_shape=None, autoencoder_latent_shape=None):
        super().__init__(scheduler)
        self.scale_factor = scale_factor
        if (ldm_latent_shape is None) ^ (autoencoder_latent_shape is None):
            raise ValueError("If ldm_latent_shape is None, autoencoder_latent_shape must be None, and vice versa.")
        if ldm_latent_shape is not None:
            raise NotImplementedError("latent resizing (ldm_latent_shape) is not used by the reference")

    @torch.no_grad()
    def sample(self, input_noise, autoencoder_model, diffusion_model, scheduler=None, save_intermediates=False,
               intermediate_steps=100, conditioning=None, mode="crossattn", verbose=True, seg=None, step_noise=None,
               seed=None, return_latent=False):
        if not isinstance(autoencoder_model, AutoencoderKL):
            raise TypeError("b200ldm inferers drive b200ldm.AutoencoderKL (there is no generic / CPU fallback)")
        latent, lat_inter, loop = super().sample(
            input_noise, diffusion_model, scheduler, save_intermediates, intermediate_steps, conditioning, mode,
            verbose, seg, step_noise=step_noise, seed=seed, _return_loop=True)
        image = self._decode(autoencoder_model, loop)
        if save_intermediates:
            inter = [autoencoder_model.decode_stage_2_outputs(l / self.scale_factor) for l in lat_inter]
            return image, inter
        if return_latent:
            return image, latent
        return image

    def _decode(self, ae: AutoencoderKL, loop: SamplingLoop) -> torch.Tensor:
        """decode(x_0 / scale_factor) straight from the device-resident NDHWC latent."""
        lat = loop.latent_ncdhw()
        if self.scale_factor != 1.0:
            inv = torch.full((lat.shape[0],), 1.0 / float(self.scale_factor), dtype=torch.float32, device=lat.device)
            zero = torch.zeros_like(inv)
            lat = ops.add_noise(lat, lat, inv, zero)  # lat * (1/scale) on our own kernel
        return ae.decode_graphed(lat) if USE_GRAPH else ae.decode_stage_2_outputs(lat)

    @torch.no_grad()
    def __call__(self, inputs, autoencoder_model, diffusion_model, noise, timesteps, condition=None, mode="crossattn",
                 seg=None, encode_noise=None):
        latent = autoencoder_model.encode_stage_2_inputs(inputs, noise=encode_noise)
        if self.scale_factor != 1.0:
            sf = torch.full((latent.shape[0],), float(self.scale_factor), dtype=torch.float32, device=latent.device)
            latent = ops.add_noise(latent, latent, sf, torch.zeros_like(sf))
        return super().__call__(latent, diffusion_model, noise, timesteps, condition, mode, seg)
