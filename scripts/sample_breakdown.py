"""Where the wall time of LatentDiffusionInferer.sample goes (setup / capture / loop / decode). usage: [T] [latent]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
import torch
import b200ldm
import bench as B
from b200ldm.inferer import SamplingLoop

T = int(sys.argv[1]) if len(sys.argv) > 1 else 200
L = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = "cuda"
unet = B.seeded_init(b200ldm.DiffusionModelUNet(**B.UNET_32G, precision="bf16"), 1234).to(dev).eval()
ae = B.seeded_init(b200ldm.AutoencoderKL(**B.AE_32G, precision="bf16"), 4321).to(dev).eval()
sched = b200ldm.DDPMScheduler(**B.SCHED); sched.set_timesteps(T)
inf = b200ldm.LatentDiffusionInferer(sched, 1.0)
x = torch.randn((8, 16, L, L, L), device=dev); c = torch.randn((8, 16, L, L, L), device=dev)


def tick(label, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"  {label:28s} {1e3 * (t1 - t0):9.1f} ms", flush=True)
    return t1


for rep in range(3):
    print(f"--- pass {rep} (T={T}, latent {L}^3, batch 8)")
    torch.cuda.synchronize(); t0 = time.perf_counter(); ts = t0
    loop = SamplingLoop(unet, sched, x, c, "concat", seed=1, use_graph=True)
    t0 = tick("SamplingLoop.__init__", t0)
    loop.step()
    t0 = tick("first step (warm-up+capture)", t0)
    for _ in range(T - 1):
        loop.step()
    t0 = tick(f"{T - 1} graph replays", t0)
    img = inf._decode(ae, loop)
    t0 = tick("decode", t0)
    print(f"  total {1e3 * (t0 - ts):.1f} ms")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    img = inf.sample(input_noise=x, autoencoder_model=ae, diffusion_model=unet, scheduler=sched, conditioning=c,
                     mode="concat", verbose=False, seed=rep)
    tick("inferer.sample (whole call)", t0)
    del loop
