"""GroupNorm apply/stats micro-benchmark (variant from B200LDM_GN_VARIANT)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
import torch
from b200ldm import ops
dev = "cuda"
for (N, S, C) in [(8, 44352, 256), (8, 44352, 512), (8, 512000, 64), (8, 5544, 512)]:
    x = torch.randn((N, S, 1, 1, C), device=dev).bfloat16()
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts, ta = [], []
    for it in range(6):
        flush.zero_()
        stats = ops.zeros_f64(N * 64, dev)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        ops._lib.check(ops._lib.load().b200ldm_groupnorm_stats(x.data_ptr(), None, stats.data_ptr(), N, S, C, 0, 32, 1, ops._stream()))
        e[1].record()
        y = ops.groupnorm(x, g, b, 32, 1e-6, True, stats=stats)
        e[2].record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(e[0].elapsed_time(e[1])); ta.append(e[1].elapsed_time(e[2]))
    nb = N * S * C * 2
    print(f"variant {os.environ.get('B200LDM_GN_VARIANT', '0')} N{N} S{S} C{C}: stats {min(ts)*1e3:7.1f} us {nb/min(ts)/1e6:6.0f} GB/s | apply {min(ta)*1e3:7.1f} us {2*nb/min(ta)/1e6:6.0f} GB/s")
