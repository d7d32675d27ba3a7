"""One eager UNet denoise step between cudaProfilerStart/Stop, for an ncu launch list of exactly that step:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
      python scripts/ncu_step.py [32g|16g] [latent edge or d,h,w] [batch]
(warm-up, the conv kernels' measured choices and weight repacking happen before the profiled range)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
import torch
import b200ldm
from b200ldm.inferer import SamplingLoop
import bench as B

cfg = sys.argv[1] if len(sys.argv) > 1 else "32g"
lat = sys.argv[2] if len(sys.argv) > 2 else "20"
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 8
lat = tuple(int(v) for v in lat.split(",")) if "," in lat else (int(lat),) * 3
kw = B.UNET_32G if cfg == "32g" else B.UNET_16G
unet = B.seeded_init(b200ldm.DiffusionModelUNet(**kw, precision="bf16"), 1234).to("cuda").eval()
sched = b200ldm.DDPMScheduler(**B.SCHED)
c_lat = kw["out_channels"]
x_T = torch.randn((batch, c_lat) + lat)
cond = torch.randn((batch, kw["in_channels"] - c_lat) + lat).cuda() if kw["in_channels"] > c_lat else None
loop = SamplingLoop(unet, sched, x_T.cuda(), cond, "concat" if cond is not None else "crossattn", use_graph=False)
for _ in range(3):
    loop._body()
torch.cuda.synchronize()
torch.cuda.profiler.start()
loop._body()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step")
