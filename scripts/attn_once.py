"""A few launches of one attention shape (for ncu captures): SHAPE=B,H,L,d python scripts/attn_once.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
import torch
from b200ldm import ops
B, H, L, d = (int(v) for v in os.environ.get("SHAPE", "8,4,5544,64").split(","))
qkv = torch.randn((B, L, 3 * H * d), device="cuda").to(torch.bfloat16)
for _ in range(4):
    ops.attention(qkv, H, d, d ** -0.5)
torch.cuda.synchronize()
print("done")
