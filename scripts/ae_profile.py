"""AutoencoderKL-32g decode (+ optional reconstruct) per-kernel breakdown: python scripts/ae_profile.py [latent edge] [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
import torch
import b200ldm
from b200ldm import ops
import bench as B

lat = (int(sys.argv[1]) if len(sys.argv) > 1 else 20,) * 3
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = "cuda"
ae = B.seeded_init(b200ldm.AutoencoderKL(**B.AE_32G, precision="bf16"), 4321).to(dev).eval()
z = ops.to_ndhwc(torch.randn((batch, 16) + lat).to(dev), 64, torch.bfloat16)
for _ in range(2):
    ae.decode_ndhwc(z)
torch.cuda.synchronize()
ops.PROFILER = ops.KernelProfiler()
ae.decode_ndhwc(z); ae.decode_ndhwc(z)
torch.cuda.synchronize()
kinds, tags = ops.PROFILER.summary()
ops.PROFILER = None
for k, v in sorted(kinds.items(), key=lambda kv: -kv[1]["ms"]):
    rate = v["work"] / (v["ms"] * 1e-3)
    print(f"   {k:10s} {v['ms'] / 2:8.3f} ms  {v['launches'] // 2:4d} launches  " + (f"{rate / 1e12:7.1f} TF/s" if k in ("conv3d", "attention") else f"{rate / 1e9:7.0f} GB/s"))
for (k, t), v in sorted(tags.items(), key=lambda kv: -kv[1]["ms"])[:12]:
    if k == "conv3d":
        print(f"      {t:38s} {v['ms'] / 2:7.3f} ms x{v['launches'] // 2:3d} {v['work'] / (v['ms'] * 1e-3) / 1e12:7.1f} TF/s")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ae.decode_ndhwc(z)
e1.record(); torch.cuda.synchronize()
print(f"decode eager: {e0.elapsed_time(e1) / 5:.3f} ms (B200LDM_CONV_HALO={os.environ.get('B200LDM_CONV_HALO', '1')})")
