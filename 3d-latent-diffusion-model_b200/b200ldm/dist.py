"""Multi-GPU driver for the sampling path: the batch of independent volumes is
partitioned across ranks (one process per GPU, full model replica each, no per-step
communication); the only collective is the final gather of decoded volumes and seeds.
The reference has no distributed inference (3d_ldm/inference.py:91 loops serially); its
training-side process-group setup is 3d_ldm/utils.py:55-63.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """(rank, world, local_rank); initialises torch.distributed when WORLD_SIZE > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) of volumes owned by `rank`; remainders go to the lowest ranks."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rank_seed(base_seed: int, rank: int, index: int = 0) -> int:
    """Per-volume seed: base + 1000*rank + local index."""
    return base_seed + 1000 * rank + index


def gather_volumes(local: torch.Tensor, seeds: torch.Tensor, total: int):
    """All-gather per-rank outputs [n_r, ...] (ragged over ranks allowed) and int64 seeds -> ([total, ...], [total])
    on every rank.  One all_gather_into_tensor each; padded to the largest shard."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local, seeds
    world = dist.get_world_size()
    n_max = (total + world - 1) // world
    pad = n_max - local.shape[0]
    if pad:
        local = torch.cat([local, local.new_zeros((pad,) + tuple(local.shape[1:]))])
        seeds = torch.cat([seeds, seeds.new_full((pad,), -1)])
    out = local.new_empty((world * n_max,) + tuple(local.shape[1:]))
    sout = seeds.new_empty((world * n_max,))
    dist.all_gather_into_tensor(out, local.contiguous())
    dist.all_gather_into_tensor(sout, seeds.contiguous())
    keep = []
    for r in range(world):
        lo, hi = shard_range(total, r, world)
        keep += list(range(r * n_max, r * n_max + (hi - lo)))
    idx = torch.as_tensor(keep, device=out.device)
    return out.index_select(0, idx), sout.index_select(0, idx)


@torch.no_grad()
def sample_sharded(inferer, autoencoder, unet, scheduler, total: int, latent_shape, base_seed: int = 42,
                   conditioning=None, mode: str = "crossattn", device=None, gather: bool = True, noise=None,
                   max_batch: int | None = None):
    """Generate `total` volumes across all ranks.  Rank r draws x_T for its shard from
    torch.Generator(cpu).manual_seed(rank_seed(base_seed, r, i)) (as inference.py draws noise on the CPU, :88-92),
    samples + decodes them - as one batch, or in chunks of `max_batch` volumes (every full chunk replays the same
    captured graphs) - then gathers.  Returns (volumes [total,...], seeds [total]).
    `noise` (optional): the x_T of ALL `total` volumes, host (pinned for an asynchronous upload) or device; the rank's
    slice is used instead of drawing.  `conditioning` likewise covers all `total` volumes on either side."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    device = device or (torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cpu")
    lo, hi = shard_range(total, rank, world)
    seeds = [rank_seed(base_seed, rank, i) for i in range(hi - lo)]
    if noise is not None:
        x_T = noise[lo:hi]
    elif seeds:
        x_T = torch.stack([torch.randn(tuple(latent_shape), generator=torch.Generator("cpu").manual_seed(s))
                           for s in seeds])
    else:
        x_T = torch.empty((0,) + tuple(latent_shape))
    step = max_batch if max_batch else max(1, hi - lo)
    parts = []
    for c0 in range(0, hi - lo, step):
        c1 = min(c0 + step, hi - lo)
        cond = None if conditioning is None else conditioning[lo + c0:lo + c1].to(device, non_blocking=True)
        parts.append(inferer.sample(input_noise=x_T[c0:c1].to(device, non_blocking=True), autoencoder_model=autoencoder,
                                    diffusion_model=unet, scheduler=scheduler, conditioning=cond, mode=mode,
                                    verbose=False, seed=seeds[c0]))
    vols = (parts[0] if len(parts) == 1 else torch.cat(parts)) if parts else torch.empty((0,), device=device)
    sd = torch.as_tensor(seeds, dtype=torch.int64, device=device)
    if not gather:
        return vols, sd
    return gather_volumes(vols, sd, total)
