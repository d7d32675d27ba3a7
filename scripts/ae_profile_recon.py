"""AutoencoderKL-32g reconstruct (encode -> mu -> decode) per-kernel breakdown at one edge: python scripts/ae_profile_recon.py [edge]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
import torch
import b200ldm
from b200ldm import ops
import bench as B

edge = int(sys.argv[1]) if len(sys.argv) > 1 else 160
dev = "cuda"
ae = B.seeded_init(b200ldm.AutoencoderKL(**B.AE_32G, precision="bf16"), 4321).to(dev).eval()
v = torch.rand((1, 1, edge, edge, edge), device=dev)
for _ in range(2):
    ae.reconstruct(v)
torch.cuda.synchronize()
ops.PROFILER = ops.KernelProfiler()
n = 3
for _ in range(n):
    ae.reconstruct(v)
torch.cuda.synchronize()
kinds, tags = ops.PROFILER.summary()
ops.PROFILER = None
tot = sum(x["ms"] for x in kinds.values()) / n
print(f"== AE reconstruct {edge}^3: profiled kernels {tot:.3f} ms")
for k, x in sorted(kinds.items(), key=lambda kv: -kv[1]["ms"]):
    rate = x["work"] / (x["ms"] * 1e-3)
    print(f"   {k:10s} {x['ms'] / n:8.3f} ms  {x['launches'] // n:4d} launches  " + (f"{rate / 1e12:7.1f} TF/s" if k in ("conv3d", "attention") else f"{rate / 1e9:7.0f} GB/s"))
for (k, t), x in sorted(tags.items(), key=lambda kv: -kv[1]["ms"])[:32]:
    rate = x["work"] / (x["ms"] * 1e-3)
    print(f"      {k:10s} {t:42s} {x['ms'] / n:7.3f} ms x{x['launches'] // n:3d} " + (f"{rate / 1e12:7.1f} TF/s" if k in ("conv3d", "attention") else f"{rate / 1e9:7.0f} GB/s"))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ae.reconstruct(v)
e1.record(); torch.cuda.synchronize()
print(f"reconstruct eager: {e0.elapsed_time(e1) / 5:.3f} ms")
