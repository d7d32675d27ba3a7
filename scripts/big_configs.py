"""Sanity + timing of the large BASELINE.json configs on one GPU (cfg 4: DDIM-50 @ 40^3 latent, cfg 5: AE reconstruct)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
import torch
import b200ldm
import bench as B
dev = "cuda"
unet = B.seeded_init(b200ldm.DiffusionModelUNet(**B.UNET_32G, precision="bf16"), 1234).to(dev).eval()
ae = B.seeded_init(b200ldm.AutoencoderKL(**B.AE_32G, precision="bf16"), 4321).to(dev).eval()

def timed(fn, n=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): out = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n, out

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "cfg4"):
    sched = b200ldm.DDIMScheduler(**B.SCHED); sched.set_timesteps(50)
    inf = b200ldm.LatentDiffusionInferer(sched, 1.0)
    x = torch.randn((4, 16, 40, 40, 40)); c = torch.randn((4, 16, 40, 40, 40))
    kw = dict(input_noise=x.to(dev), autoencoder_model=ae, diffusion_model=unet, scheduler=sched, conditioning=c.to(dev), mode="concat", verbose=False)
    t, img = timed(lambda: inf.sample(**kw))
    t, img = timed(lambda: inf.sample(**kw))
    print(f"cfg4: DDIM-50 @40^3 latent, batch 4 -> {tuple(img.shape)}: {t:.3f} s = {4 / t:.2f} volumes/s; finite={bool(torch.isfinite(img).all())}; "
          f"algorithmic {(50 * 31.64 + 4 * 21.74) / t:.0f} TFLOP/s; mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
if which in ("all", "cfg5"):
    for edge in (96, 128, 160):
        v = torch.rand((1, 1, edge, edge, edge), device=dev)
        t, rec = timed(lambda: ae.reconstruct(v))
        t, rec = timed(lambda: ae.reconstruct(v), 2)
        fl = {96: 1.75 + 3.27, 128: 5.42 + 9.03, 160: 14.68 + 21.74}[edge]
        print(f"cfg5: AE reconstruct {edge}^3: {t * 1e3:.1f} ms = {1 / t:.2f} volumes/s, {fl / t:.0f} TFLOP/s algorithmic; finite={bool(torch.isfinite(rec).all())}; "
              f"mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
