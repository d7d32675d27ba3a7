import os, sys
sys.path.insert(0, "/root/repo/3d-latent-diffusion-model_b200")
import torch
from b200ldm import ops
for B, H, L, d in [(2, 1, 8000, 256), (8, 1, 8000, 256), (8, 1, 1000, 256), (8, 2, 300, 128)]:
    qkv = torch.randn((B, L, 3 * H * d), device="cuda").to(torch.bfloat16)
    o = ops.attention(qkv, H, d, d ** -0.5)
    torch.cuda.synchronize()
    ref = ops.attention(qkv, H, d, d ** -0.5, engine=ops.ENGINE_SIMT) if L <= 1000 else None
    torch.cuda.synchronize()
    print(B, H, L, d, "ok", float(o.float().abs().mean()), None if ref is None else float((o.float() - ref.float()).abs().max()), flush=True)
