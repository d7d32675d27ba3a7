"""Stress the AutoencoderKL decode (and head-dim-256 attention) for intermittent failures: python scripts/stress_decode.py [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
import torch
import b200ldm
import bench as B
from b200ldm import ops

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = "cuda"
ae = B.seeded_init(b200ldm.AutoencoderKL(**B.AE_32G, precision="bf16"), 4321).to(dev).eval()
z = torch.randn((8, 16, 20, 20, 20), device=dev)
ref = ae.decode(z)
torch.cuda.synchronize()
print("first decode ok", float(ref.abs().mean()), flush=True)
for i in range(iters):
    out = ae.decode(z)
    torch.cuda.synchronize()
    d = float((out - ref).abs().max())
    if not (d < 0.05):
        print("iter", i, "MISMATCH", d, flush=True)
qkv = torch.randn((8, 8000, 768), device=dev).to(torch.bfloat16)
o0 = ops.attention(qkv, 1, 256, 256 ** -0.5)
for i in range(iters * 3):
    o = ops.attention(qkv, 1, 256, 256 ** -0.5)
    torch.cuda.synchronize()
    if not torch.equal(o, o0):
        print("attention iter", i, "differs", float((o.float() - o0.float()).abs().max()), flush=True)
print("stress done", flush=True)
