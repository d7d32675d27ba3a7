"""Time the tcgen05 flash-attention kernel alone (CUDA events, L2 flushed between launches).
usage: python scripts/attn_bench.py   -> one line per (shape, V encoding)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
import torch
from b200ldm import ops

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
SHAPES = [(8, 4, 5544, 64), (8, 8, 693, 64), (8, 8, 1000, 64), (4, 8, 8000, 64), (4, 16, 1000, 64), (1, 1, 8000, 256),
          (1, 1, 64000, 256)]
for B, H, L, d in SHAPES:
    qkv = torch.randn((B, L, 3 * H * d), device=dev).to(torch.bfloat16)
    for _ in range(1):
        t = qkv
        for _ in range(2):
            ops.attention(t, H, d, d ** -0.5)
        ms = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.attention(t, H, d, d ** -0.5); e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        best = min(ms)
        print(f"attention B{B} H{H} L{L} d{d}: {best:.3f} ms  "
              f"{4.0 * B * H * L * L * d / best / 1e9:.0f} TFLOP/s", flush=True)
