"""Per-kernel CUDA-event breakdown of one UNet denoise step and one AutoencoderKL decode (eager, KernelProfiler).
usage: python scripts/profile_models.py [32g|16g] [latent edge or d,h,w] [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
import torch
import b200ldm
from b200ldm import ops
from b200ldm.inferer import SamplingLoop
import bench as B

cfg = sys.argv[1] if len(sys.argv) > 1 else "32g"
lat = sys.argv[2] if len(sys.argv) > 2 else "20"
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 8
lat = tuple(int(v) for v in lat.split(",")) if "," in lat else (int(lat),) * 3
dev = "cuda"
kw = B.UNET_32G if cfg == "32g" else B.UNET_16G
unet = B.seeded_init(b200ldm.DiffusionModelUNet(**kw, precision="bf16"), 1234).to(dev).eval()
sched = b200ldm.DDPMScheduler(**B.SCHED)
c_lat = kw["out_channels"]
x_T = torch.randn((batch, c_lat) + lat)
cond = torch.randn((batch, kw["in_channels"] - c_lat) + lat).to(dev) if kw["in_channels"] > c_lat else None
loop = SamplingLoop(unet, sched, x_T.to(dev), cond, "concat" if cond is not None else "crossattn", use_graph=True,
                    batch_streams=int(os.environ.get("PROFILE_STREAMS", "1")))


def report(title, fn, n=2):
    fn(); torch.cuda.synchronize()
    ops.PROFILER = ops.KernelProfiler()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    kinds, tags = ops.PROFILER.summary()
    dump = os.environ.get("TAGDUMP")
    if dump:
        for i, (kind, tag, work, e0, e1) in enumerate(ops.PROFILER.records):
            if dump in tag:
                prev = ops.PROFILER.records[i - 1]
                print(f"   [{i}] {kind} {tag}: {e0.elapsed_time(e1):.3f} ms  (prev: {prev[0]} {prev[1]} {prev[3].elapsed_time(prev[4]):.3f} ms; gap {prev[4].elapsed_time(e0):.3f} ms)")
    ops.PROFILER = None
    tot = sum(v["ms"] for v in kinds.values()) / n
    print(f"== {title}: profiled kernels {tot:.3f} ms")
    for k, v in sorted(kinds.items(), key=lambda kv: -kv[1]["ms"]):
        rate = v["work"] / (v["ms"] * 1e-3)
        unit = f"{rate / 1e12:7.1f} TF/s" if k in ("conv3d", "attention") else f"{rate / 1e9:7.0f} GB/s"
        print(f"   {k:10s} {v['ms'] / n:8.3f} ms  {v['launches'] // n:4d} launches  {unit}")
    for (k, t), v in sorted(tags.items(), key=lambda kv: -kv[1]["ms"])[:int(os.environ.get("TOPN", 14))]:
        rate = v["work"] / (v["ms"] * 1e-3)
        unit = f"{rate / 1e12:7.1f} TF/s" if k in ("conv3d", "attention") else f"{rate / 1e9:7.0f} GB/s"
        print(f"      {k:9s} {t:38s} {v['ms'] / n:7.3f} ms x{v['launches'] // n:3d} {unit}")


report(f"UNet-{cfg} step, batch {batch}, latent {lat}", loop._body)
for _ in range(3):
    loop.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    loop.step()
e1.record(); torch.cuda.synchronize()
print(f"graph replay: {e0.elapsed_time(e1) / 10:.3f} ms/step, {loop.launches_per_step} launches/step")
if cfg == "32g":
    ae = B.seeded_init(b200ldm.AutoencoderKL(**B.AE_32G, precision="bf16"), 4321).to(dev).eval()
    z = ops.to_ndhwc(torch.randn((batch, 16) + lat).to(dev), 64, torch.bfloat16)
    report(f"AE-32g decode, batch {batch}, latent {lat} -> x4", lambda: ae.decode_ndhwc(z))
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        ae.decode_ndhwc(z)
    e1.record(); torch.cuda.synchronize()
    print(f"decode eager: {e0.elapsed_time(e1) / 3:.3f} ms")
