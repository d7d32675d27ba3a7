#!/usr/bin/env python
"""Pin the oracle (and through it every parity number of this repo) to real MONAI, the day MONAI is importable.

    python scripts/verify_against_monai.py            exit 0 = oracle == MONAI, 1 = mismatch, 3 = MONAI not importable

The reference's arithmetic lives in un-vendored `monai>=1.3.0` (/root/reference/3d_ldm/requirements.txt:7; the imports
at 3d_ldm/inference.py:23-24 need >= 1.4); it is not installable in the build container, so `oracle/` was written from
the MONAI specification in SURVEY.md App. A-E and DESIGN.md says "parity unpinned".  This script is the pin: with MONAI
present it checks, class by class,

  1. state_dict key lists and shapes (oracle vs monai.networks.nets.{DiffusionModelUNet, AutoencoderKL}) for the
     reference's configs (config_train_32g / 16g / multigpu / stable; meta device, no memory) - and that a MONAI
     state_dict loads strictly into the oracle AND into the b200ldm modules;
  2. one forward per class on identical random (non-zero) weights and inputs: UNet noise prediction, AutoencoderKL
     encode / decode - max relative error <= 1e-5;
  3. one `step` per scheduler (DDPM, DDIM eta 0 / 0.5, several timesteps, same generator seed) and `add_noise`;
  4. `LatentDiffusionInferer.sample` for 5 steps with concat conditioning (same torch.manual_seed on both sides).

`tests/test_oracle.py::test_oracle_matches_monai_when_importable` runs it under pytest.importorskip("monai").
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "3d-latent-diffusion-model_b200"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

UNETS = {
    "config_train_32g": dict(spatial_dims=3, in_channels=32, out_channels=16, channels=[256, 512, 1024],
                             attention_levels=[False, True, True], num_head_channels=[0, 64, 64], num_res_blocks=2),
    "config_train_16g": dict(spatial_dims=3, in_channels=8, out_channels=8, channels=[256, 256, 512],
                             attention_levels=[False, True, True], num_head_channels=[0, 64, 64], num_res_blocks=2),
    "config_train_stable": dict(spatial_dims=3, in_channels=16, out_channels=8, channels=[128, 256, 512],
                                attention_levels=[False, False, True], num_head_channels=[0, 0, 32], num_res_blocks=1),
}
AES = {
    "config_train_32g": dict(spatial_dims=3, in_channels=1, out_channels=1, latent_channels=16, channels=[64, 128, 256],
                             num_res_blocks=2, norm_num_groups=32, norm_eps=1e-6, attention_levels=[False, False, True],
                             with_encoder_nonlocal_attn=False, with_decoder_nonlocal_attn=False),
    "config_train_multigpu": dict(spatial_dims=3, in_channels=1, out_channels=1, latent_channels=16,
                                  channels=[64, 128, 256], num_res_blocks=2, norm_num_groups=32, norm_eps=1e-6,
                                  attention_levels=[False, True, True], with_encoder_nonlocal_attn=False,
                                  with_decoder_nonlocal_attn=True),
    "config_train_stable": dict(spatial_dims=3, in_channels=1, out_channels=1, latent_channels=8, channels=[32, 64, 128],
                                num_res_blocks=1, norm_num_groups=16, norm_eps=1e-6,
                                attention_levels=[False, False, True], with_encoder_nonlocal_attn=False,
                                with_decoder_nonlocal_attn=False),
}
SCHED = dict(num_train_timesteps=1000, schedule="scaled_linear_beta", beta_start=0.0015, beta_end=0.0195)


def relerr(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


class Report:
    def __init__(self):
        self.failures = []

    def check(self, ok: bool, what: str, detail=""):
        print(("ok   " if ok else "FAIL ") + what + (f"  [{detail}]" if detail else ""))
        if not ok:
            self.failures.append(what)


def verify(monai, rep: Report) -> None:
    import b200ldm
    import golden_cases as gc
    import oracle
    from monai.inferers import LatentDiffusionInferer as MInferer
    from monai.networks.nets import AutoencoderKL as MAE
    from monai.networks.nets import DiffusionModelUNet as MUNet
    from monai.networks.schedulers import DDIMScheduler as MDDIM
    from monai.networks.schedulers import DDPMScheduler as MDDPM

    # 1. key lists / shapes on the reference's configs (meta device: no storage)
    for kind, table, mcls, ocls, pcls in (("unet", UNETS, MUNet, oracle.DiffusionModelUNet, b200ldm.DiffusionModelUNet),
                                          ("ae", AES, MAE, oracle.AutoencoderKL, b200ldm.AutoencoderKL)):
        for name, kw in table.items():
            with torch.device("meta"):
                m, o = mcls(**kw), ocls(**kw)
            ms, os_ = m.state_dict(), o.state_dict()
            only_m, only_o = sorted(set(ms) - set(os_)), sorted(set(os_) - set(ms))
            rep.check(not only_m and not only_o, f"{kind} {name}: state_dict keys equal",
                      f"only MONAI: {only_m[:4]} only oracle: {only_o[:4]}")
            bad = [k for k in ms if k in os_ and ms[k].shape != os_[k].shape]
            rep.check(not bad, f"{kind} {name}: parameter shapes equal", str(bad[:4]))
            try:
                with torch.device("meta"):
                    p = pcls(**kw)
                rep.check(set(p.state_dict()) == set(ms), f"{kind} {name}: b200ldm state_dict keys equal MONAI's")
            except NotImplementedError as e:
                rep.check(False, f"{kind} {name}: b200ldm instantiates", str(e))

    # 2. forwards on the tiny seeded cases (identical non-zero weights)
    g = torch.Generator().manual_seed(0)
    m_unet = oracle.randomize_parameters(MUNet(**gc.UNET_KW), 1234).eval()
    o_unet = oracle.DiffusionModelUNet(**gc.UNET_KW).eval()
    o_unet.load_state_dict(m_unet.state_dict())
    x = torch.randn((2, gc.UNET_KW["in_channels"]) + gc.LATENT, generator=g)
    ts = torch.tensor([37.0, 912.0])
    with torch.no_grad():
        e = relerr(o_unet(x, ts), m_unet(x, ts))
    rep.check(e <= 1e-5, "DiffusionModelUNet.forward oracle vs MONAI", f"{e:.2e}")
    m_x = oracle.randomize_parameters(MUNet(**gc.XATTN_UNET_KW), 2468).eval()
    o_x = oracle.DiffusionModelUNet(**gc.XATTN_UNET_KW).eval()
    o_x.load_state_dict(m_x.state_dict())
    xx = torch.randn((2, 4) + gc.LATENT, generator=g)
    ctx = torch.randn((2, gc.XATTN_CTX_LEN, 24), generator=g)
    lab = torch.tensor([3, 0])
    with torch.no_grad():
        e = relerr(o_x(xx, ts, context=ctx, class_labels=lab), m_x(xx, ts, context=ctx, class_labels=lab))
    rep.check(e <= 1e-5, "DiffusionModelUNet.forward (cross-attention, class embeddings) oracle vs MONAI", f"{e:.2e}")
    m_ae = oracle.randomize_parameters(MAE(**gc.AE_KW), 4321).eval()
    o_ae = oracle.AutoencoderKL(**gc.AE_KW).eval()
    o_ae.load_state_dict(m_ae.state_dict())
    z = torch.randn((2, gc.AE_KW["latent_channels"]) + gc.LATENT, generator=g)
    img = torch.rand((1, 1) + tuple(2 * v for v in gc.LATENT), generator=g)
    with torch.no_grad():
        e = relerr(o_ae.decode_stage_2_outputs(z), m_ae.decode_stage_2_outputs(z))
        rep.check(e <= 1e-5, "AutoencoderKL.decode_stage_2_outputs oracle vs MONAI", f"{e:.2e}")
        (mu, sg), (mmu, msg) = o_ae.encode(img), m_ae.encode(img)
        e = max(relerr(mu, mmu), relerr(sg, msg))
        rep.check(e <= 1e-5, "AutoencoderKL.encode oracle vs MONAI", f"{e:.2e}")

    # 3. schedulers
    for n in (1000, 10):
        ms, os_ = MDDPM(**SCHED), oracle.DDPMScheduler(**SCHED)
        ms.set_timesteps(n)
        os_.set_timesteps(n)
        rep.check(torch.equal(torch.as_tensor(ms.timesteps), torch.as_tensor(os_.timesteps)), f"DDPM timesteps n={n}")
        xs, eps = torch.randn((2, 4) + gc.LATENT, generator=g), torch.randn((2, 4) + gc.LATENT, generator=g)
        for t in (int(ms.timesteps[0]), int(ms.timesteps[n // 2]), 0):
            a = ms.step(eps, t, xs, generator=torch.Generator().manual_seed(5))
            b = os_.step(eps, t, xs, generator=torch.Generator().manual_seed(5))
            e = max(relerr(b[0], a[0]), relerr(b[1], a[1]))
            rep.check(e <= 1e-6, f"DDPMScheduler.step n={n} t={t}", f"{e:.2e}")
    ms, os_ = MDDIM(**SCHED), oracle.DDIMScheduler(**SCHED)
    ms.set_timesteps(50)
    os_.set_timesteps(50)
    rep.check(torch.equal(torch.as_tensor(ms.timesteps), torch.as_tensor(os_.timesteps)), "DDIM timesteps n=50")
    for eta in (0.0, 0.5):
        for t in (980, 500, 0):
            a = ms.step(eps, t, xs, eta=eta, generator=torch.Generator().manual_seed(6))
            b = os_.step(eps, t, xs, eta=eta, generator=torch.Generator().manual_seed(6))
            e = max(relerr(b[0], a[0]), relerr(b[1], a[1]))
            rep.check(e <= 1e-6, f"DDIMScheduler.step eta={eta} t={t}", f"{e:.2e}")
    tt = torch.tensor([0, 999])
    e = relerr(os_.add_noise(xs, eps, tt), ms.add_noise(xs, eps, tt))
    rep.check(e <= 1e-6, "add_noise", f"{e:.2e}")

    # 4. the sampling call of the reference (3d_ldm/train_diffusion.py:326-333), 5 steps
    ms, os_ = MDDPM(**SCHED), oracle.DDPMScheduler(**SCHED)
    ms.set_timesteps(5)
    os_.set_timesteps(5)
    x_T, cond, _ = gc.sampling_inputs(5)
    torch.manual_seed(77)
    a = MInferer(ms, scale_factor=1.0).sample(input_noise=x_T, autoencoder_model=m_ae, diffusion_model=m_unet,
                                              scheduler=ms, conditioning=cond, mode="concat", verbose=False)
    torch.manual_seed(77)
    b = oracle.LatentDiffusionInferer(os_, 1.0).sample(x_T, o_ae, o_unet, os_, conditioning=cond, mode="concat",
                                                       verbose=False)
    e = relerr(b, a)
    rep.check(e <= 1e-4, "LatentDiffusionInferer.sample (DDPM, 5 steps, concat) oracle vs MONAI", f"{e:.2e}")


def main() -> int:
    try:
        import monai
    except Exception as e:  # noqa: BLE001
        print(f"MONAI is not importable here ({type(e).__name__}: {e}) - the oracle stays unpinned")
        return 3
    print(f"MONAI {monai.__version__}")
    rep = Report()
    verify(monai, rep)
    if rep.failures:
        print(f"{len(rep.failures)} mismatch(es) between the oracle and MONAI:")
        for f in rep.failures:
            print("  " + f)
        return 1
    print("oracle == MONAI on every check: parity is pinned to this MONAI version")
    return 0


if __name__ == "__main__":
    sys.exit(main())
