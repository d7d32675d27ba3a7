"""A/B of the sub-batch stream split of the denoising step (B200LDM_BATCH_STREAMS; since commit "top-level split" only
the TOP resolution level of the UNet is split, before that the whole step): graph-replayed ms/step for the 32g UNet at
20^3 (cfg 3) and the 16g UNet at 36x44x28 (cfg 2), batch 8, alternating 1 / 2 / 4 streams in one process.
usage: python scripts/ab_streams.py [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
import torch
import b200ldm
from b200ldm.inferer import SamplingLoop
import bench as B

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sched = b200ldm.DDPMScheduler(**B.SCHED)
for cfg, kw, lat in (("32g", B.UNET_32G, (20, 20, 20)), ("16g", B.UNET_16G, B.LATENT_16G)):
    unet = B.seeded_init(b200ldm.DiffusionModelUNet(**kw, precision="bf16"), 1234).to("cuda").eval()
    c_lat = kw["out_channels"]
    x_T = torch.randn((8, c_lat) + lat)
    cond = torch.randn((8, kw["in_channels"] - c_lat) + lat).cuda() if kw["in_channels"] > c_lat else None
    mode = "concat" if cond is not None else "crossattn"
    loops = {n: SamplingLoop(unet, sched, x_T.cuda(), cond, mode, use_graph=True, batch_streams=n) for n in (1, 2, 4)}
    for n, loop in loops.items():
        for _ in range(4):
            loop.step()
    torch.cuda.synchronize()
    for r in range(reps):
        for n, loop in loops.items():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                loop.step()
            e1.record(); torch.cuda.synchronize()
            print(f"{cfg} streams={n}: {e0.elapsed_time(e1) / 20:.3f} ms/step ({loop.launches_per_step} launches)", flush=True)
    del loops, unet
    torch.cuda.empty_cache()
