"""Parity at the sizes and trajectory lengths BASELINE.json names, bf16 tensor-core path
against the fp32 CPU oracle on identical weights / x_T / conditioning / per-step noise (tests/long_cases.py):

  cfg 3  DDPM, all 1000 steps, 32g UNet + AutoencoderKL, 20^3 latent -> 80^3 volume: against the committed fixture
         tests/golden/cfg3_ddpm1000_20.pt (the oracle's latents after 10 / 100 / 500 / 1000 steps + its decoded volume;
         written by `scripts/long_parity.py cfg3 --fixture ...` on the GPU box's host cores - 531 s of CPU oracle, which is
         why this test does not recompute it).
  cfg 5  32g AutoencoderKL reconstruct at 96^3 (oracle recomputed live: 3 s), 128^3 behind B200LDM_SLOW=1.
  cfg 4  DDIM-50 at a 40^3 latent -> 160^3 (5 min of CPU oracle): behind B200LDM_SLOW=1; the measured run is
         profiles/r02_long_parity/lp_cfg4_40.log.
Gate: decoded volume PSNR >= 40 dB with peak = max(ref) - min(ref)."""
import os

import pytest
import torch

import long_cases as lc
import oracle

pytestmark = pytest.mark.gpu
SLOW = os.environ.get("B200LDM_SLOW", "0") != "0"
FIX = os.path.join(os.path.dirname(__file__), "golden", "cfg3_ddpm1000_20.pt")


def test_cfg3_ddpm_1000_steps_vs_oracle_fixture():
    import b200ldm
    fx = torch.load(FIX, weights_only=True)
    assert (fx["kind"], fx["latent"], fx["n_steps"]) == ("ddpm", 20, 1000)
    r_unet, r_ae = lc.oracle_models()
    x_T, cond, noise = lc.sampling_inputs(20, 1000)
    unet = lc.port(r_unet, b200ldm.DiffusionModelUNet, lc.UNET_32G, "bf16")
    ae = lc.port(r_ae, b200ldm.AutoencoderKL, lc.AE_32G, "bf16")
    lat, inter = lc.sample_gpu(unet, "ddpm", 1000, x_T, cond, noise, fx["every"])
    img = ae.decode_stage_2_outputs(lat)
    steps = lc.saved_steps(lc.scheduler(oracle, "ddpm", 1000), fx["every"])
    got = {s: lc.psnr(inter[steps.index(s)], ref) for s, ref in fx["latents"].items()}
    p_img = lc.psnr(img, fx["image"])
    print("cfg3 DDPM-1000 @ 20^3, bf16 vs oracle fixture: latent PSNR by step",
          {s: round(v, 1) for s, v in got.items()}, f"decoded volume PSNR {p_img:.1f} dB")
    assert img.shape == (1, 1, 80, 80, 80)
    assert p_img >= 40.0, p_img
    assert got[1000] >= 40.0 and got[10] >= 80.0 and got[100] >= 70.0 and got[500] >= 55.0, got


@pytest.mark.parametrize("edge", [96, pytest.param(128, marks=pytest.mark.skipif(not SLOW, reason="B200LDM_SLOW=1"))])
def test_cfg5_autoencoder_reconstruct_vs_oracle(edge):
    """closes SURVEY.md 8a5 at the 32g widths: 3-level encoder (two AEDownsample, d = 256 attention at (edge/4)^3
    tokens) + decoder, against the oracle"""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    r = lc.run_cfg5(edge)
    print(r)
    assert r["psnr_reconstruct_db"] >= 40.0 and r["psnr_mu_db"] >= 40.0 and r["psnr_decode_of_oracle_latent_db"] >= 40.0, r
    assert r["relerr_mu"] <= 2e-2 and r["relerr_reconstruct"] <= 4e-2, r


@pytest.mark.skipif(not SLOW, reason="5 min of CPU oracle: B200LDM_SLOW=1 (measured: profiles/r02_long_parity/lp_cfg4_40.log)")
def test_cfg4_ddim_50_steps_40cubed_vs_oracle():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    r = lc.run_sampling_vs_oracle("ddim", 40, 50, every=10)
    print(r)
    assert r["psnr_image_db"] >= 40.0, r


@pytest.mark.skipif(not SLOW, reason="7 min of the fp32 validation engine: B200LDM_SLOW=1")
def test_cfg3_full_batch_bf16_vs_fp32_engine():
    r = lc.run_sampling_bf16_vs_fp32("ddpm", 20, 1000, 8, every=100)
    print(r)
    assert min(r["psnr_image_per_volume_db"]) >= 40.0, r
