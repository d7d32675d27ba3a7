"""Which phase bounds the persistent conv kernel on a narrow conv?  Times 64 -> 64 (and 128 -> 128) k3 convs with the
kernel pinned (B200LDM_CONV_KERNEL=2) and the debug modes of B200LDM_CONV_ASHIFT: 0 normal, 1 no activation loads,
2 no weight loads, 3 no loads, 4 no epilogue, 7 MMA issue only.  Results of modes != 0 are garbage; timing only."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
import torch
from b200ldm import ops
dev, dt = "cuda", torch.bfloat16
for C, S in ((64, 80), (128, 40)):
    x = torch.randn((8, S, S, S, C), device=dev).to(dt)
    w = ops.repack_conv_weight((torch.randn(C, C, 3, 3, 3) * (C * 27) ** -0.5).to(dev), C, C, dt)
    for _ in range(2):
        ops.conv3d(x, w, None, C, 3, 1, 1, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.conv3d(x, w, None, C, 3, 1, 1, 1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"ASHIFT={os.environ.get('B200LDM_CONV_ASHIFT', '0')} {C}->{C} k3 @8x{S}^3: {ms:.3f} ms "
          f"({2.0 * 8 * S ** 3 * C * C * 27 / ms / 1e9:.0f} TFLOP/s nominal)", flush=True)
