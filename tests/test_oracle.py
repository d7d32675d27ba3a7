"""Oracle self-tests (CPU).  The reference pins nothing on this path, so these are the
pins we create: closed-form scheduler anchors (App. C.4), parameter counts, state-dict key snapshot,
and the committed golden vectors."""
import hashlib
import os

import pytest
import torch

import golden_cases as gc
import oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "tiny_cases.pt")
CFG32 = dict(spatial_dims=3, in_channels=32, out_channels=16, channels=[256, 512, 1024],
             attention_levels=[False, True, True], num_head_channels=[0, 64, 64], num_res_blocks=2)
AE32 = dict(spatial_dims=3, in_channels=1, out_channels=1, latent_channels=16, channels=[64, 128, 256],
            num_res_blocks=2, norm_num_groups=32, norm_eps=1e-6, attention_levels=[False, False, True],
            with_encoder_nonlocal_attn=False, with_decoder_nonlocal_attn=False)


def test_scheduler_anchors():
    s = oracle.DDPMScheduler(1000, "scaled_linear_beta", beta_start=0.0015, beta_end=0.0195)
    assert abs(s.betas[0].item() - 0.0015) < 1e-9 and abs(s.betas[999].item() - 0.0195) < 1e-8
    assert abs(s.betas[1].item() - 0.00150783) < 1e-8
    assert abs(s.alphas_cumprod[0].item() - 0.99850) < 1e-6
    assert abs(s.alphas_cumprod[499].item() - 0.115844) < 1e-6
    assert abs(s.alphas_cumprod[999].item() - 1.42304e-4) < 1e-9
    s.set_timesteps(10)
    assert s.timesteps.tolist() == list(range(900, -1, -100))
    d = oracle.DDIMScheduler(1000, "scaled_linear_beta", beta_start=0.0015, beta_end=0.0195)
    d.set_timesteps(50)
    assert d.timesteps.tolist() == list(range(980, -1, -20))


def test_ddpm_step_closed_form():
    s = oracle.DDPMScheduler(1000, "scaled_linear_beta", beta_start=0.0015, beta_end=0.0195)
    g = torch.Generator().manual_seed(0)
    x, e, nz = (torch.randn(1, 4, 3, 3, 3, generator=g) for _ in range(3))
    prev, x0 = s.step(e, 999, x, noise=nz)
    a, ap, b = s.alphas_cumprod[999].double(), s.alphas_cumprod[998].double(), s.betas[999].double()
    x0_ref = ((x.double() - (1 - a).sqrt() * e.double()) / a.sqrt()).clamp(-1, 1)
    c0, ct, var = ap.sqrt() * b / (1 - a), (1 - b).sqrt() * (1 - ap) / (1 - a), (1 - ap) / (1 - a) * b
    assert abs(c0.item() - 2.34953e-4) < 1e-8 and abs(ct.item() - 0.990199) < 1e-5 and abs(var.item() - 0.0194999) < 1e-6
    ref = c0 * x0_ref + ct * x.double() + var.sqrt() * nz.double()
    assert torch.allclose(prev.double(), ref, atol=2e-5) and torch.allclose(x0.double(), x0_ref, atol=1e-4)
    prev0, _ = s.step(e, 0, x)  # t == 0: no noise, alpha_prev = 1
    a0 = s.alphas_cumprod[0].double()
    x0_0 = ((x.double() - (1 - a0).sqrt() * e.double()) / a0.sqrt()).clamp(-1, 1)
    assert torch.allclose(prev0.double(), x0_0, atol=1e-5)


def test_ddim_step_is_deterministic_and_final_alpha_one():
    d = oracle.DDIMScheduler(1000, "scaled_linear_beta", beta_start=0.0015, beta_end=0.0195)
    d.set_timesteps(50)
    g = torch.Generator().manual_seed(1)
    x, e = torch.randn(1, 4, 2, 2, 2, generator=g), torch.randn(1, 4, 2, 2, 2, generator=g)
    p1, x0 = d.step(e, 0, x)
    p2, _ = d.step(e, 0, x)
    assert torch.equal(p1, p2)
    assert torch.allclose(p1, x0, atol=1e-6)  # alpha_prev = 1 -> x_prev = x0


def test_add_noise_matches_definition():
    s = oracle.DDPMScheduler(1000, "scaled_linear_beta", beta_start=0.0015, beta_end=0.0195)
    g = torch.Generator().manual_seed(2)
    x0, nz = torch.randn(3, 2, 2, 2, 2, generator=g), torch.randn(3, 2, 2, 2, 2, generator=g)
    t = torch.tensor([0, 499, 999])
    out = s.add_noise(x0, nz, t)
    for i, ti in enumerate(t.tolist()):
        a = s.alphas_cumprod[ti]
        assert torch.allclose(out[i], a.sqrt() * x0[i] + (1 - a).sqrt() * nz[i], atol=1e-6)


def test_param_counts_and_keys():
    u = oracle.DiffusionModelUNet(**CFG32)
    n = sum(p.numel() for p in u.parameters())
    assert n == 666_065_936, n
    ae = oracle.AutoencoderKL(**AE32)
    ne, nd = (sum(p.numel() for p in m.parameters()) for m in (ae.encoder, ae.decoder))
    assert (ne, nd) == (9_423_568, 12_741_889), (ne, nd)
    keys = sorted(u.state_dict().keys()) + sorted(ae.state_dict().keys())
    digest = hashlib.sha256("\n".join(keys).encode()).hexdigest()
    # snapshot of the MONAI (>=1.4) key names of SURVEY.md App. E; a change here breaks checkpoint loading
    assert len(u.state_dict()) == 320
    for k in ("conv_in.conv.weight", "time_embed.2.bias", "down_blocks.0.downsampler.op.conv.weight",
              "down_blocks.1.attentions.0.attn.to_q.weight", "middle_block.attention.attn.out_proj.bias",
              "up_blocks.0.upsampler.postconv.conv.weight", "up_blocks.2.resnets.2.skip_connection.conv.bias",
              "out.0.weight", "out.2.conv.bias"):
        assert k in u.state_dict(), k
    for k in ("encoder.blocks.0.conv.weight", "encoder.blocks.3.conv.conv.weight", "decoder.blocks.2.attn.to_v.bias",
              "decoder.blocks.5.postconv.conv.weight", "decoder.blocks.6.nin_shortcut.conv.weight", "decoder.blocks.8.postconv.conv.bias",
              "decoder.blocks.11.weight", "decoder.blocks.12.conv.bias", "quant_conv_mu.conv.weight",
              "quant_conv_log_sigma.conv.bias", "post_quant_conv.conv.weight"):
        assert k in ae.state_dict(), k
    assert digest == KEY_DIGEST, digest


KEY_DIGEST = "711750ec5644778929fcdfd2db273c19ea7840f04b1f6d159fa712b087889f4c"


def test_random_init_unet_predicts_zero_and_reinit_does_not():
    u = oracle.DiffusionModelUNet(**gc.UNET_KW).eval()
    x = torch.randn((1, 8) + gc.LATENT)
    with torch.no_grad():
        assert u(x, torch.tensor([5.0])).abs().max().item() == 0.0  # zero-initialised out conv (SURVEY.md gotcha 1)
        oracle.randomize_parameters(u, 1)
        assert u(x, torch.tensor([5.0])).abs().max().item() > 1e-2


def test_timestep_embedding_cos_first():
    from oracle.nets import get_timestep_embedding
    e = get_timestep_embedding(torch.tensor([0.0, 3.0]), 8)
    assert torch.allclose(e[0], torch.tensor([1, 1, 1, 1, 0, 0, 0, 0.0]))
    assert abs(e[1, 0].item() - torch.cos(torch.tensor(3.0)).item()) < 1e-6
    assert abs(e[1, 4].item() - torch.sin(torch.tensor(3.0)).item()) < 1e-6


def test_context_without_conditioning_raises():
    u = oracle.DiffusionModelUNet(**gc.UNET_KW)
    with pytest.raises(ValueError):
        u(torch.zeros((1, 8) + gc.LATENT), torch.tensor([1.0]), context=torch.zeros(1, 1, 4))


def test_golden_vectors():
    gold = torch.load(GOLD, weights_only=True)
    with torch.no_grad():
        unet, x, ts = gc.tiny_unet_case()
        assert torch.allclose(unet(x, ts), gold["unet_eps"], atol=2e-5, rtol=1e-4)
        ae, z = gc.tiny_ae_case()
        assert torch.allclose(ae.decode_stage_2_outputs(z), gold["ae_decode"], atol=5e-5, rtol=1e-4)
        mu, sigma = ae.encode(gc.tiny_ae_image())
        assert torch.allclose(mu, gold["ae_mu"], atol=5e-5, rtol=1e-4)
        assert torch.allclose(sigma, gold["ae_sigma"], atol=5e-5, rtol=1e-4)
        for kind in ("ddpm", "ddim"):
            img, lat = gc.tiny_sampling_case(kind)
            assert torch.allclose(lat, gold[f"sample_{kind}_latent"], atol=2e-4, rtol=1e-3)
            assert torch.allclose(img, gold[f"sample_{kind}_image"], atol=2e-3, rtol=1e-3)


def test_inferer_concat_requires_conditioning():
    s = gc.make_scheduler(oracle, "ddpm", 2)
    inf = oracle.LatentDiffusionInferer(s)
    with pytest.raises(ValueError):
        inf.sample(torch.zeros((1, 4) + gc.LATENT), gc.tiny_ae(), gc.tiny_unet(), s, mode="concat")
    with pytest.raises(NotImplementedError):
        inf.sample(torch.zeros((1, 4) + gc.LATENT), gc.tiny_ae(), gc.tiny_unet(), s, mode="nope")


XGOLD = os.path.join(os.path.dirname(__file__), "golden", "xattn_cases.pt")


def test_cross_attention_golden_and_context_sensitivity():
    """with_conditioning=True (SpatialTransformer + class embedding): the oracle equals its committed
    fixture, the context and the class label both move the prediction, and a context the attention cannot tell apart
    (every token identical) gives the same result for any context length (softmax over equal keys = the value)."""
    gold = torch.load(XGOLD, weights_only=True)
    with torch.no_grad():
        unet, x, ts, ctx, labels = gc.tiny_xattn_case()
        eps = unet(x, ts, context=ctx, class_labels=labels)
        assert torch.allclose(eps, gold["xattn_unet_eps"], atol=2e-5, rtol=1e-4)
        assert (unet(x, ts, context=ctx.flip(0), class_labels=labels) - eps).abs().max() > 1e-3
        assert (unet(x, ts, context=ctx, class_labels=labels.flip(0)) - eps).abs().max() > 1e-3
        one = ctx[:, :1]
        a = unet(x, ts, context=one, class_labels=labels)
        b = unet(x, ts, context=one.expand(-1, 5, -1).contiguous(), class_labels=labels)
        assert torch.allclose(a, b, atol=1e-5, rtol=1e-5)
        _, img, lat = gc.tiny_xattn_sampling_case("ddpm")
        assert torch.allclose(lat, gold["xattn_sample_latent"], atol=2e-4, rtol=1e-3)
        assert torch.allclose(img, gold["xattn_sample_image"], atol=2e-3, rtol=1e-3)


def test_cross_attention_constructor_surface():
    with pytest.raises(ValueError):
        oracle.DiffusionModelUNet(**{**gc.XATTN_UNET_KW, "cross_attention_dim": None})
    with pytest.raises(ValueError):
        oracle.DiffusionModelUNet(**{**gc.UNET_KW, "cross_attention_dim": 8})
    u = gc.tiny_xattn_unet()
    with pytest.raises(ValueError):  # num_class_embeds set, labels missing
        u(torch.zeros((1, 4) + gc.LATENT), torch.tensor([1.0]), context=torch.zeros(1, 2, 24))
    keys = [k for k in u.state_dict() if k.startswith("middle_block.attention.")]
    assert "middle_block.attention.transformer_blocks.1.attn2.to_k.weight" in keys
    assert "middle_block.attention.transformer_blocks.0.attn1.to_q.bias" not in keys       # qkv_bias=False
    assert u.state_dict()["middle_block.attention.transformer_blocks.0.attn2.to_k.weight"].shape == (128, 24)
    assert u.state_dict()["middle_block.attention.transformer_blocks.0.ff.linear1.weight"].shape == (1024, 128)
    assert u.state_dict()["class_embedding.weight"].shape == (5, 256)


def _load_verify_script():
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "verify_against_monai.py")
    spec = importlib.util.spec_from_file_location("verify_against_monai", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_oracle_matches_monai_when_importable():
    """THE pin of this repo's parity chain: skipped while MONAI is not installable;
    with MONAI present every key list, forward, scheduler step and a 5-step sample must agree."""
    monai = pytest.importorskip("monai")
    v = _load_verify_script()
    rep = v.Report()
    v.verify(monai, rep)
    assert not rep.failures, rep.failures


def test_verify_script_logic_with_stand_in(monkeypatch):
    """The verification script itself must not rot while MONAI is absent: run it against a stand-in `monai` package that
    re-exports the oracle's classes (says nothing about MONAI - only that every check executes and passes on equal
    implementations, and that a deliberately different implementation is caught)."""
    import sys
    import types
    v = _load_verify_script()
    pk = {n: types.ModuleType(n) for n in ("monai", "monai.inferers", "monai.networks", "monai.networks.nets",
                                           "monai.networks.schedulers")}
    pk["monai"].__version__ = "stand-in"
    pk["monai.inferers"].LatentDiffusionInferer = oracle.LatentDiffusionInferer
    pk["monai.networks.nets"].AutoencoderKL = oracle.AutoencoderKL
    pk["monai.networks.nets"].DiffusionModelUNet = oracle.DiffusionModelUNet
    pk["monai.networks.schedulers"].DDPMScheduler = oracle.DDPMScheduler
    pk["monai.networks.schedulers"].DDIMScheduler = oracle.DDIMScheduler
    for n, m in pk.items():
        monkeypatch.setitem(sys.modules, n, m)
    rep = v.Report()
    v.verify(pk["monai"], rep)
    assert not rep.failures, rep.failures

    class SkewedDDPM(oracle.DDPMScheduler):      # a scheduler that differs must be reported
        def step(self, *a, **k):
            prev, x0 = super().step(*a, **k)
            return prev * 1.001, x0

    pk["monai.networks.schedulers"].DDPMScheduler = SkewedDDPM
    rep = v.Report()
    v.verify(pk["monai"], rep)
    assert any("DDPMScheduler.step" in f for f in rep.failures)
