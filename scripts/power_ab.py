"""Is the replayed UNet step power-bound?  Graph-replays the step for a few seconds and reports ms/step next to the
nvidia-smi SM clock / board power sampled DURING the replays:  python scripts/power_ab.py [16g|32g] [steps]
(run once per setting of B200LDM_CONV_STREAMK / other switches; they are read when the library loads)."""
import os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
import torch
import b200ldm
from b200ldm.inferer import SamplingLoop
import bench as B

cfg = sys.argv[1] if len(sys.argv) > 1 else "16g"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
kw, lat = (B.UNET_16G, B.LATENT_16G) if cfg == "16g" else (B.UNET_32G, (20, 20, 20))
unet = B.seeded_init(b200ldm.DiffusionModelUNet(**kw, precision="bf16"), 1234).to("cuda").eval()
sched = b200ldm.DDPMScheduler(**B.SCHED)
c_lat = kw["out_channels"]
x_T = torch.randn((8, c_lat) + lat)
cond = torch.randn((8, kw["in_channels"] - c_lat) + lat).cuda() if kw["in_channels"] > c_lat else None
loop = SamplingLoop(unet, sched, x_T.cuda(), cond, "concat" if cond is not None else "crossattn", use_graph=True)
for _ in range(5):
    loop.step()
torch.cuda.synchronize()
for label, n in (("first 10 replays", 10), (f"{steps} replays", steps)):
    with B.ClockSampler(0) as clk:
        ms = B.timed_region(loop.step, n, 1)
    rows = clk.rows
    pw = [float(r[3]) for r in rows if len(r) > 3 and r[3].replace(".", "", 1).isdigit()]
    sm = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "", 1).isdigit()]
    print(f"{cfg} {label}: {ms / n:.3f} ms/step; SM clock median {statistics.median(sm) if sm else None} MHz, "
          f"board power median {statistics.median(pw) if pw else None} W ({len(sm)} samples); {clk.summary()['reasons']}; "
          f"STREAMK={os.environ.get('B200LDM_CONV_STREAMK', '1')}")
