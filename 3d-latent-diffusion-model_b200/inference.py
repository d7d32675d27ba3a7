#!/usr/bin/env python
"""Drop-in for the reference CLI `3d_ldm/inference.py` on B200 (same flags, same JSON configs, same control flow):

    python inference.py -e ./config/environment.json -c ./config/config_train_32g.json -n 4 [--steps 1000] [--random-init]

Differences from the reference, all forced by its own inconsistencies:
  * the UNet is declared with in_channels = 2 x latent (concat conditioning, 3d_ldm/train_diffusion.py:197-205) while
    inference.py:88-89 feeds latent-channel noise only -> a conditioning latent is required: `--condition file.pt`
    (a [1,C,d,h,w] tensor) or, without it, a seeded N(0,1) latent (reported on stdout);
  * `--random-init` replaces the two torch.load() calls when no checkpoints exist (smoke runs);
  * volumes are written with the built-in NIfTI-1 writer (nibabel is an optional dependency of the reference only).
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import sys
from datetime import datetime
from pathlib import Path

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from b200ldm import DDPMScheduler, LatentDiffusionInferer, define_instance  # noqa: E402
from b200ldm.nifti import save_nifti  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description="B200 Latent Diffusion Model Inference")
    ap.add_argument("-e", "--environment-file", default="./config/environment.json")
    ap.add_argument("-c", "--config-file", default="./config/config_train_32g.json")
    ap.add_argument("-n", "--num", type=int, default=1, help="number of generated images")
    ap.add_argument("--steps", type=int, default=None, help="DDPM inference steps (default: num_train_timesteps)")
    ap.add_argument("--condition", default=None, help=".pt file with the conditioning latent (concat mode)")
    ap.add_argument("--random-init", action="store_true", help="seeded random weights instead of checkpoints")
    ap.add_argument("--seed", type=int, default=42)
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("b200ldm needs a B200 (there is no CPU fallback)")
    device = torch.device("cuda")
    for f in (args.environment_file, args.config_file):
        for k, v in json.load(open(f)).items():
            setattr(args, k, v)
    torch.manual_seed(args.seed)

    autoencoder = define_instance(args, "autoencoder_def")
    unet = define_instance(args, "diffusion_def")
    if args.random_init:
        g = torch.Generator().manual_seed(args.seed)
        with torch.no_grad():
            for m in (autoencoder, unet):
                for p in m.parameters():
                    p.copy_(torch.randn(p.shape, generator=g) / max(1.0, float(p[0].numel()) ** 0.5) if p.ndim > 1
                            else (1.0 + 0.1 * torch.randn(p.shape, generator=g)))
    else:
        autoencoder.load_state_dict(torch.load(os.path.join(args.model_dir, "autoencoder.pt"), weights_only=True))
        unet.load_state_dict(torch.load(os.path.join(args.model_dir, "diffusion_unet.pt"), weights_only=True))
    autoencoder, unet = autoencoder.to(device).eval(), unet.to(device).eval()

    ns = args.NoiseScheduler
    scheduler = DDPMScheduler(num_train_timesteps=ns["num_train_timesteps"], schedule="scaled_linear_beta",
                              beta_start=ns["beta_start"], beta_end=ns["beta_end"])
    if args.steps:
        scheduler.set_timesteps(args.steps)
    inferer = LatentDiffusionInferer(scheduler, scale_factor=1.0)

    out_dir = getattr(args, "output_dir", "./output")
    Path(out_dir).mkdir(parents=True, exist_ok=True)
    latent_shape = [p // 4 for p in args.diffusion_train["patch_size"]]
    noise_shape = [1, args.latent_channels] + latent_shape
    n_cond = unet.in_channels - args.latent_channels
    cond = None
    if n_cond > 0:
        if args.condition:
            cond = torch.load(args.condition, weights_only=True).float()
        else:
            cond = torch.randn([1, n_cond] + latent_shape, generator=torch.Generator().manual_seed(args.seed + 1))
            logging.info("no --condition given: using a seeded N(0,1) conditioning latent %s", tuple(cond.shape))
    for i in range(args.num):
        noise = torch.randn(noise_shape, dtype=torch.float32).to(device)   # CPU draw then move, as the reference
        with torch.no_grad():
            img = inferer.sample(input_noise=noise, autoencoder_model=autoencoder, diffusion_model=unet,
                                 scheduler=scheduler, conditioning=None if cond is None else cond.to(device),
                                 mode="concat" if cond is not None else "crossattn", verbose=False)
        name = os.path.join(out_dir, datetime.now().strftime("synimg_%Y%m%d_%H%M%S") + f"_{i}")
        path = save_nifti(img[0, 0].unsqueeze(-1).cpu().numpy(), name)
        logging.info("wrote %s  shape %s", path, tuple(img.shape))


if __name__ == "__main__":
    logging.basicConfig(stream=sys.stdout, level=logging.INFO,
                        format="[%(asctime)s.%(msecs)03d][%(levelname)5s](%(name)s) - %(message)s",
                        datefmt="%Y-%m-%d %H:%M:%S")
    main()
