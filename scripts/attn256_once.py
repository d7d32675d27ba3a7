import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
import torch
from b200ldm import ops
qkv = torch.randn((1, 64000, 768), device="cuda").to(torch.bfloat16)
for _ in range(3):
    ops.attention(qkv, 1, 256, 256 ** -0.5)
torch.cuda.synchronize()
