"""CPU tests of the host side: C-ABI export list, config resolver (the drop-in boundary), module surface /
state-dict keys vs the oracle, schedulers' coefficient tables vs the oracle's step, batch sharding and the
world_size-2 gather over gloo.  No kernel is launched here."""
import ctypes
import json
import os
import re
import socket
import sys

import pytest
import torch

import golden_cases as gc
import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol(lib_built):
    hdr = open(os.path.join(ROOT, "include", "b200ldm.h")).read()
    declared = set(re.findall(r"\b(b200ldm_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(lib_built)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200ldm.h but not exported"
    from b200ldm import _lib
    assert set(_lib.SIGNATURES) == declared
    assert _lib.load().b200ldm_version() == 100


def test_struct_layout_matches_header(lib_built):
    from b200ldm import _lib
    # the compiled structs are the reference: 8 pointers + 20 int32 + workspace ptr + int64 + chunk_stats ptr +
    # 3 skip pointers + 2 int32 (conv), 7 pointers + 2 int64 + 4 int32 + uint64 + seed pointer + 2 learned-variance
    # pointers + int64 Philox offset (sched)
    lib = _lib.load()
    assert ctypes.sizeof(_lib.ConvParams) == lib.b200ldm_sizeof_params(0) == 8 * 8 + 20 * 4 + 8 + 8 + 8 + 3 * 8 + 2 * 4 + 2 * 4   # + chunk_width + defer_finalize
    assert ctypes.sizeof(_lib.SchedParams) == lib.b200ldm_sizeof_params(1) == 7 * 8 + 2 * 8 + 4 * 4 + 8 + 8 + 2 * 8 + 8


def test_no_cpu_fallback(lib_built):
    import b200ldm
    from b200ldm import ops
    with pytest.raises(RuntimeError):
        ops.silu_f32(torch.zeros(4))
    u = b200ldm.DiffusionModelUNet(**gc.UNET_KW)
    s = gc.make_scheduler(b200ldm, "ddpm", 3)
    with pytest.raises(RuntimeError):
        b200ldm.DiffusionInferer(s).sample(torch.zeros((1, 8) + gc.LATENT), u, s, verbose=False)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "3d-latent-diffusion-model_b200", "b200ldm")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(import|from)\s+oracle", src, re.M), f


@pytest.mark.parametrize("cfg", ["config_train_32g.json", "config_train_16g.json"])
def test_define_instance_on_reference_style_config(cfg):
    import b200ldm
    # the reference's own JSON (copied shape: 3d_ldm/config/*.json) - keys this path reads
    conf = {
        "spatial_dims": 3, "image_channels": 1, "latent_channels": 16 if "32g" in cfg else 8,
        "autoencoder_def": {"_target_": "monai.networks.nets.AutoencoderKL" if "32g" in cfg else "networks.AutoencoderKL",
                            "spatial_dims": "@spatial_dims", "in_channels": "$@image_channels",
                            "out_channels": "@image_channels", "latent_channels": "@latent_channels",
                            "channels": [64, 128, 256], "num_res_blocks": 2, "norm_num_groups": 32, "norm_eps": 1e-06,
                            "attention_levels": [False, False, True], "with_encoder_nonlocal_attn": False,
                            "with_decoder_nonlocal_attn": False},
        "diffusion_def": {"_target_": "monai.networks.nets.DiffusionModelUNet", "spatial_dims": "@spatial_dims",
                          "in_channels": "$@latent_channels * 2", "out_channels": "@latent_channels",
                          "channels": [256, 512], "attention_levels": [False, True], "num_head_channels": [0, 64],
                          "num_res_blocks": 1},
    }

    class Args:
        pass

    a = Args()
    for k, v in conf.items():
        setattr(a, k, v)
    ae = b200ldm.define_instance(a, "autoencoder_def")
    un = b200ldm.define_instance(a, "diffusion_def")
    assert isinstance(ae, b200ldm.AutoencoderKL) and ae.latent_channels == conf["latent_channels"]
    assert isinstance(un, b200ldm.DiffusionModelUNet) and un.in_channels == 2 * conf["latent_channels"]


def test_state_dict_keys_and_shapes_match_oracle():
    import b200ldm
    for ours, ref in ((b200ldm.DiffusionModelUNet(**gc.UNET_KW), oracle.DiffusionModelUNet(**gc.UNET_KW)),
                      (b200ldm.DiffusionModelUNet(**gc.XATTN_UNET_KW), oracle.DiffusionModelUNet(**gc.XATTN_UNET_KW)),
                      (b200ldm.AutoencoderKL(**gc.AE_KW), oracle.AutoencoderKL(**gc.AE_KW))):
        a, b = ours.state_dict(), ref.state_dict()
        assert list(a.keys()) == list(b.keys())
        assert all(a[k].shape == b[k].shape for k in a)
        ours.load_state_dict(ref.state_dict())  # strict


def test_upsampler_key_spellings_both_load_strict():
    """MONAI >= 1.4 names the UNet's post-upsample conv `up_blocks.{i}.upsampler.postconv.conv.*`;
    SURVEY.md App. E recalled `...upsampler.upsample.postconv.conv.*`.  state_dict() emits the former, both load strictly
    (product and oracle), and a key that is neither still fails."""
    import b200ldm
    for mod in (b200ldm, oracle):
        ref = oracle.randomize_parameters(oracle.DiffusionModelUNet(**gc.UNET_KW), 3)
        sd = ref.state_dict()
        flat = [k for k in sd if ".upsampler." in k]
        assert flat and all(".upsampler.postconv.conv." in k for k in flat), flat
        nested = {k.replace(".upsampler.postconv.", ".upsampler.upsample.postconv."): v for k, v in sd.items()}
        assert any(".upsampler.upsample.postconv." in k for k in nested)
        for variant in (sd, nested):
            m = mod.DiffusionModelUNet(**gc.UNET_KW)
            m.load_state_dict(variant)  # strict
            for k in flat:
                assert torch.equal(m.state_dict()[k], sd[k])
        bad = {k.replace(".upsampler.postconv.", ".upsampler.wrong.postconv."): v for k, v in sd.items()}
        with pytest.raises(RuntimeError):
            mod.DiffusionModelUNet(**gc.UNET_KW).load_state_dict(bad)


def test_load_old_state_dict_generative_format():
    """checkpoints of the pre-1.4 `generative` package: attention linears directly on the block with `proj_attn`,
    transformer `to_out.0`, post-upsample conv named `conv` (MONAI's load_old_state_dict surface)."""
    import b200ldm

    def to_old(sd):
        out = {}
        for k, v in sd.items():
            if "transformer_blocks" in k:
                k = k.replace(".out_proj.", ".to_out.0.")
            else:
                k = k.replace(".attn.out_proj.", ".proj_attn.").replace(".attn.to_", ".to_")
            out[k.replace("postconv", "conv")] = v
        return out

    for cls, kw, seed in ((b200ldm.DiffusionModelUNet, gc.UNET_KW, 5), (b200ldm.DiffusionModelUNet, gc.XATTN_UNET_KW, 6),
                          (b200ldm.AutoencoderKL, gc.AE_KW, 7)):
        ref = oracle.randomize_parameters(getattr(oracle, cls.__name__)(**kw), seed)
        sd = ref.state_dict()
        old = to_old(sd)
        assert set(old) != set(sd)
        m = cls(**kw)
        m.load_old_state_dict(old)
        got = m.state_dict()
        assert list(got) == list(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
        m2 = cls(**kw)
        m2.load_old_state_dict(sd)   # already in the new format: plain load
        assert all(torch.equal(m2.state_dict()[k], sd[k]) for k in sd)
        broken = dict(old)
        broken.pop(next(k for k in broken if "proj_attn" in k or "to_out" in k))
        with pytest.raises(RuntimeError):
            cls(**kw).load_old_state_dict(broken)


def test_constructor_errors_match_reference_surface():
    import b200ldm
    with pytest.raises(ValueError):
        b200ldm.DiffusionModelUNet(3, 8, 4, channels=(64, 100), attention_levels=(False, True), num_head_channels=64)
    with pytest.raises(ValueError):
        b200ldm.DiffusionModelUNet(3, 8, 4, channels=(64, 128), attention_levels=(False,), num_head_channels=64)
    with pytest.raises(ValueError):
        b200ldm.AutoencoderKL(3, channels=(64, 128), attention_levels=(False,))
    u = b200ldm.DiffusionModelUNet(**gc.UNET_KW)
    with pytest.raises(ValueError):
        u(torch.zeros((1, 8) + gc.LATENT), torch.tensor([1.0]), context=torch.zeros(1, 1, 4))
    with pytest.raises(ValueError):   # with_conditioning without a context width / a width without with_conditioning
        b200ldm.DiffusionModelUNet(**{**gc.XATTN_UNET_KW, "cross_attention_dim": None})
    with pytest.raises(ValueError):
        b200ldm.DiffusionModelUNet(**{**gc.UNET_KW, "cross_attention_dim": 8})
    xu = b200ldm.DiffusionModelUNet(**gc.XATTN_UNET_KW)
    with pytest.raises(ValueError):   # class embeddings configured, labels missing
        xu(torch.zeros((1, 4) + gc.LATENT), torch.tensor([1.0]), context=torch.zeros(1, 2, 24))


@pytest.mark.parametrize("pred", ["epsilon", "sample", "v_prediction"])
@pytest.mark.parametrize("kind,n", [("ddpm", 10), ("ddpm", 1000), ("ddim", 50)])
def test_coef_table_reproduces_oracle_step(kind, n, pred):
    """x0 = clamp(a_x*(x + r*m) | r*m); x' = c_x0*x0 + c_xt*x + c_e*m + sigma*z with our host-side table == oracle.step,
    for every prediction type and (DDIM) eta in {0, 0.7}."""
    import b200ldm
    ours, ref = gc.make_scheduler(b200ldm, kind, n, prediction_type=pred), gc.make_scheduler(oracle, kind, n, prediction_type=pred)
    assert ours.timesteps.tolist() == ref.timesteps.tolist()
    g = torch.Generator().manual_seed(3)
    x, e, z = (torch.randn(1, 4, 3, 3, 3, generator=g) for _ in range(3))
    for eta in ((0.0, 0.7) if kind == "ddim" else (None,)):
        for t in [int(ref.timesteps[0]), int(ref.timesteps[len(ref.timesteps) // 2]), int(ref.timesteps[-1])]:
            a_x, r, c0, ct, ce, sg, lo, hi = ours.coef_row(t) if eta is None else ours.coef_row(t, eta)
            x0 = ((x + r * e) * a_x if a_x != 0 else r * e).clamp(lo, hi)
            mine = c0 * x0 + ct * x + ce * e + sg * z
            want, want0 = ref.step(e, t, x, noise=z) if eta is None else ref.step(e, t, x, eta=eta, noise=z)
            assert torch.allclose(mine, want, atol=5e-5, rtol=2e-5), (kind, t, pred, eta)
            assert torch.allclose(x0, want0, atol=5e-5, rtol=2e-5)


@pytest.mark.parametrize("vt", ["learned", "learned_range"])
def test_learned_variance_rows_reproduce_oracle_step(vt):
    """variance_type learned / learned_range: the per-element noise scale sqrt(A*v + B) of our
    host-side rows == the oracle's step on a model output with twice the channels; without a predicted variance both
    fall through to the unclamped posterior variance."""
    import b200ldm
    ours, ref = gc.make_scheduler(b200ldm, "ddpm", 10, variance_type=vt), gc.make_scheduler(oracle, "ddpm", 10, variance_type=vt)
    assert ours.learned_variance
    g = torch.Generator().manual_seed(5)
    x, e, z = (torch.randn(1, 4, 3, 3, 3, generator=g) for _ in range(3))
    v = torch.rand(1, 4, 3, 3, 3, generator=g) * (1.6 if vt == "learned_range" else 0.02) - (0.8 if vt == "learned_range" else 0.0)
    for t in (900, 400, 0):
        a_x, r, c0, ct, ce, sg, lo, hi = ours.coef_row(t)
        A, B = ours.var_coef_row(t)
        x0 = ((x + r * e) * a_x).clamp(lo, hi)
        want, want0 = ref.step(torch.cat([e, v], dim=1), t, x, noise=z)
        mine = c0 * x0 + ct * x + (0.0 if sg == 0 else (A * v + B) ** 0.5) * z
        assert torch.allclose(mine, want, atol=5e-5, rtol=2e-5) and torch.allclose(x0, want0, atol=5e-5, rtol=2e-5)
        want_plain, _ = ref.step(e, t, x, noise=z)       # no predicted variance: unclamped fixed-small value
        assert torch.allclose(c0 * x0 + ct * x + sg * z, want_plain, atol=5e-5, rtol=2e-5)
    with pytest.raises(ValueError):
        b200ldm.DDPMScheduler(variance_type="nope")


def test_scheduler_constructor_errors():
    import b200ldm
    with pytest.raises(ValueError):
        b200ldm.DDPMScheduler(prediction_type="nonsense")
    with pytest.raises(ValueError):
        b200ldm.DDIMScheduler(prediction_type="nonsense")
    assert b200ldm.DDPMScheduler(variance_type="learned_range").learned_variance
    with pytest.raises(ValueError):
        b200ldm.DDPMScheduler(variance_type="learned_rnage")


def test_ddim_steps_offset_check_and_coef_cache_keys():
    """upstream's `steps_offset >= num_train_timesteps // num_inference_steps` ValueError (it fires from the
    constructor's set_timesteps(T) for any offset > 0, for product and oracle alike); coefficient caches follow
    attributes reassigned after construction."""
    import b200ldm
    for mod in (b200ldm, oracle):
        with pytest.raises(ValueError):
            mod.DDIMScheduler(num_train_timesteps=1000, steps_offset=1)
        s = mod.DDIMScheduler(num_train_timesteps=1000)
        s.steps_offset = 20
        with pytest.raises(ValueError):
            s.set_timesteps(50)
        s.steps_offset = 19
        s.set_timesteps(50)
        assert int(s.timesteps[0]) == 999 and int(s.timesteps[-1]) == 19
    s = b200ldm.DDPMScheduler(num_train_timesteps=1000, schedule="scaled_linear_beta", beta_start=0.0015, beta_end=0.0195)
    s.set_timesteps(10)
    k0 = s._state_key()
    t0 = s.coef_table("cpu")
    assert s.coef_table("cpu") is t0                       # memoised per (state, grid)
    s.clip_sample_values = [-2.0, 2.0]
    assert s._state_key() != k0
    t1 = s.coef_table("cpu")
    assert t1 is not t0 and float(t1[0, 6]) == -2.0 and float(t0[0, 6]) == -1.0
    s.clip_sample = False
    assert float(s.coef_table("cpu")[0, 7]) > 1e38
    s.set_timesteps(20)
    assert s.coef_table("cpu").shape[0] == 20


def test_serving_batch_plan_and_surface():
    """request batching of the serving shim (reference loop: api_server.py:163-176, one sample at a time)"""
    import b200ldm
    from b200ldm.serving import plan_batches
    assert plan_batches(0, 8) == []
    assert plan_batches(3, 8) == [(0, 3)]
    assert plan_batches(19, 8) == [(0, 8), (8, 8), (16, 3)]
    with pytest.raises(ValueError):
        plan_batches(4, 0)
    gen = b200ldm.LatentDiffusionGenerator(b200ldm.AutoencoderKL(**gc.AE_KW), b200ldm.DiffusionModelUNet(**gc.UNET_KW),
                                           dict(num_train_timesteps=1000, schedule="scaled_linear_beta",
                                                beta_start=0.0015, beta_end=0.0195), sampler="ddim", max_batch=2)
    assert gen.cond_channels == gc.UNET_KW["in_channels"] - gc.AE_KW["latent_channels"]
    with pytest.raises(RuntimeError):          # no CPU fallback
        gen.generate(torch.zeros(1, 1, 16, 16, 16), num_inference_steps=2)
    with pytest.raises(ValueError):
        gen._latent_spatial((10, 16, 16))


def test_shard_range_partitions_exactly():
    from b200ldm.dist import shard_range
    for total in (1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
    import torch.distributed as dist
    from b200ldm.dist import gather_volumes, rank_seed, shard_range
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    local = torch.stack([torch.full((1, 2, 2, 2), float(i)) for i in range(lo, hi)]) if hi > lo else torch.empty(0, 1, 2, 2, 2)
    seeds = torch.tensor([rank_seed(42, rank, i) for i in range(hi - lo)], dtype=torch.int64)
    vols, sd = gather_volumes(local, seeds, total)
    q.put((rank, vols[:, 0, 0, 0, 0].tolist(), sd.tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [4, 5])
def test_gather_world2_gloo(total):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_gather_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    from b200ldm.dist import rank_seed, shard_range
    want_seeds = []
    for r in range(2):
        lo, hi = shard_range(total, r, 2)
        want_seeds += [rank_seed(42, r, i) for i in range(hi - lo)]
    for _, vals, seeds in res:
        assert vals == [float(i) for i in range(total)]
        assert seeds == want_seeds


def test_nifti_roundtrip(tmp_path):
    """NIfTI-1 egress of the inference CLI (3d_ldm/inference.py:100-102 writes volume[..., None] with an identity affine)."""
    import numpy as np
    from b200ldm.nifti import load_nifti, save_nifti
    vol = np.random.default_rng(0).random((5, 6, 7, 1)).astype(np.float32)
    path = save_nifti(vol, str(tmp_path / "synimg_test"))
    assert path.endswith(".nii") and os.path.getsize(path) == 352 + vol.size * 4
    raw = open(path, "rb").read()
    assert raw[344:348] == b"n+1\x00" and int.from_bytes(raw[0:4], "little") == 348
    back, aff = load_nifti(path)
    assert np.array_equal(back, vol) and np.array_equal(aff, np.eye(4))
    # first index fastest on disk (what nibabel writes for a C-ordered array of this shape)
    first = np.frombuffer(raw, dtype="<f4", count=5, offset=352)
    assert np.array_equal(first, vol[:, 0, 0, 0])


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the driver's reference arm): rank 0 prints ONE JSON line with the same metric / unit /
    config keys as the GPU arm plus impl / cpu_baseline / a zero-copy e2e block; other ranks exit 0 without output."""
    import json
    import subprocess
    import sys
    bench = os.path.join(ROOT, "bench.py")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env["B200LDM_BENCH_REF_BUDGET_S"] = "6"      # keep the CPU suite short: one (cropped) volume instead of the batch
    r = subprocess.run([sys.executable, bench, "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "unet_denoise_volume_steps_per_sec"
    assert d["unit"] == "volume-steps/s" and d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["config"]["workload"].startswith("unet_step_16g")
    sys.path.insert(0, ROOT)
    import bench as bench_mod
    assert d["config"] == bench_mod.config_16g(1)          # the GPU arm's config, key for key (driver: same_config)
    # FLOP census used for the executed-vs-algorithmic fractions reproduces SURVEY.md 8d within 0.2 %
    assert abs(bench_mod.unet_flops(bench_mod.UNET_16G, bench_mod.LATENT_16G) / 2969.7e9 - 1) < 2e-3
    assert abs(bench_mod.unet_flops(bench_mod.UNET_32G, (20, 20, 20)) / 914.5e9 - 1) < 2e-3
    assert abs(bench_mod.decoder_flops(bench_mod.AE_32G, (20, 20, 20)) / 1799.7e9 - 1) < 2e-3
    assert 0.93 < bench_mod.executed_fraction_unet(bench_mod.UNET_16G, bench_mod.LATENT_16G) < 0.95
    # ... and the AutoencoderKL sweep of configs[4] (SURVEY.md 8d cfg 5: enc 1.75 / 5.42 / 14.7 / 35.9 TF, dec 3.27 / 9.03 /
    # 21.7 / 48.1 TF per volume at 96^3 / 128^3 / 160^3 / 192^3), the other sizes of the UNet and the decoder
    for edge, enc, dec in ((96, 1.75e12, 3.27e12), (128, 5.42e12, 9.03e12), (160, 14.7e12, 21.7e12), (192, 35.9e12, 48.1e12)):
        assert abs(bench_mod.encoder_flops(bench_mod.AE_32G, (edge,) * 3) / enc - 1) < 5e-3
        assert abs(bench_mod.decoder_flops(bench_mod.AE_32G, (edge // 4,) * 3) / dec - 1) < 5e-3
    for lat, f in bench_mod.FLOP_UNET_32G.items():
        assert abs(bench_mod.unet_flops(bench_mod.UNET_32G, (lat,) * 3) / f - 1) < 2e-3
    for lat, f in bench_mod.FLOP_DEC_32G.items():
        assert abs(bench_mod.decoder_flops(bench_mod.AE_32G, (lat,) * 3) / f - 1) < 2e-3
    r1 = subprocess.run([sys.executable, bench, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                        capture_output=True, text=True, env={**env, "RANK": "1", "WORLD_SIZE": "2"}, timeout=600)
    assert r1.returncode == 0 and not [l for l in r1.stdout.splitlines() if l.startswith("{")]


def test_public_header_is_plain_c():
    """include/b200ldm.h is the drop-in boundary: it must compile as C99 and as C++ on its own (no torch / CUDA types)."""
    import shutil
    import subprocess
    hdr = os.path.join(ROOT, "include", "b200ldm.h")
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                ["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    text = open(hdr).read()
    assert "#include <torch" not in text and "at::" not in text and "c10::" not in text and "cuda_runtime" not in text
