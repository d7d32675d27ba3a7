"""1x1x1 projections (QKV / out_proj GEMMs) of the UNets: persistent kernel (2) vs its wide-epilogue variant (4)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
import torch
from b200ldm import ops
dev, dt = "cuda", torch.bfloat16
for M, K, N in ((8000, 512, 1536), (8000, 512, 512), (1000, 1024, 3072), (44352, 256, 768), (44352, 256, 256), (64000, 256, 768)):
    x = torch.randn((M, K), device=dev).to(dt)
    w = ops.repack_conv_weight((torch.randn(N, K) * K ** -0.5).to(dev), N, K, dt)
    b = torch.zeros(N, device=dev)
    res = torch.randn((M, N), device=dev).to(dt) if N == K else None
    for _ in range(2):
        ops.linear(x, w, b, N, residual=res)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.linear(x, w, b, N, residual=res)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"kernel {os.environ.get('B200LDM_CONV_KERNEL', 'auto')}: [{M} x {K}] -> {N}{' +res' if res is not None else ''}: "
          f"{ms * 1e3:.1f} us ({2.0 * M * K * N / ms / 1e9:.0f} TFLOP/s)", flush=True)
