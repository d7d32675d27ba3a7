"""Parity on the configurations BASELINE.json names, at their real lengths and sizes.

    python scripts/long_parity.py cfg3 [--latent 20] [--steps 1000] [--fixture PATH]   DDPM-1000, 32g UNet + AutoencoderKL
    python scripts/long_parity.py cfg4 [--latent 40]                                   DDIM-50 at a 40^3 latent -> 160^3
    python scripts/long_parity.py cfg5 [--edge 96]                                     AutoencoderKL reconstruct (encode -> mu -> decode)
    python scripts/long_parity.py cfg3 --batch-fp32 8 --steps 1000                     bf16 engine vs the fp32 validation engine, whole batch

One volume, identical weights (oracle.randomize_parameters seeds 1234 / 4321, loaded through load_state_dict), identical
x_T / conditioning / per-step noise (torch CPU generator, seed 42); the CPU oracle (fp32, all host cores) against the bf16
tensor-core path, as the reference runs it (3d_ldm/inference.py:79-99: DDPMScheduler(1000, scaled_linear_beta,
0.0015, 0.0195), inferer.sample; 3d_ldm/train_diffusion.py:326-333: conditioning + mode="concat").

Reports the PSNR of the latent every `--every` steps and of the decoded volume; prints one JSON object and (with
--fixture) stores the oracle's checkpoint latents + decoded volume so `tests/test_gpu_long.py` can re-check the GPU path
in seconds without re-running ten minutes of CPU oracle.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "3d-latent-diffusion-model_b200"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

import long_cases as lc  # noqa: E402  (tests/long_cases.py: seeded models + inputs shared with the tests)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["cfg3", "cfg4", "cfg5"])
    ap.add_argument("--latent", type=int, default=None)
    ap.add_argument("--edge", type=int, default=96)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--every", type=int, default=10)
    ap.add_argument("--fixture", default=None)
    ap.add_argument("--batch-fp32", type=int, default=0, help="compare bf16 vs the fp32 validation engine on this batch")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    if a.which == "cfg5":
        res = lc.run_cfg5(a.edge)
    elif a.batch_fp32:
        res = lc.run_sampling_bf16_vs_fp32("ddpm" if a.which == "cfg3" else "ddim", a.latent or 20,
                                           a.steps or 1000, a.batch_fp32, a.every)
    else:
        kind = "ddpm" if a.which == "cfg3" else "ddim"
        latent = a.latent or (20 if a.which == "cfg3" else 40)
        steps = a.steps or (1000 if a.which == "cfg3" else 50)
        res = lc.run_sampling_vs_oracle(kind, latent, steps, a.every, a.fixture)
    res["host_cores"] = os.cpu_count()
    res["gpu"] = torch.cuda.get_device_name(0)
    line = json.dumps(res)
    print(line)
    if a.out:
        with open(a.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    t0 = time.time()
    main()
    print(f"# total {time.time() - t0:.0f} s", file=sys.stderr)
