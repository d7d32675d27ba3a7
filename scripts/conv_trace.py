"""Per-CTA timeline of the tcgen05 conv kernel (clock64 stamps): where do the non-MMA cycles go?"""
import ctypes, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
import torch
from b200ldm import _lib, ops

lib = _lib.load()
lib.b200ldm_debug_set_conv_trace.argtypes = [ctypes.c_void_p]
dev = "cuda"
D, H, W = (int(v) for v in os.environ.get("DHW", "36,44,28").split(","))
N, C, Cout = 8, int(os.environ.get("CIN", 256)), int(os.environ.get("COUT", 256))
x = torch.randn((N, D, H, W, C), device=dev).bfloat16()
w = ops.repack_conv_weight(torch.randn((Cout, C, 3, 3, 3), device=dev) / math.sqrt(C * 27), Cout, C, torch.bfloat16)
b = torch.zeros(Cout, device=dev)
for _ in range(3):
    y = ops.conv3d(x, w, b, Cout)
torch.cuda.synchronize()
trace = torch.zeros((4096, 8), dtype=torch.int64, device=dev)
lib.b200ldm_debug_set_conv_trace(trace.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); y = ops.conv3d(x, w, b, Cout); e1.record()
torch.cuda.synchronize()
lib.b200ldm_debug_set_conv_trace(None)
t = trace.cpu()
t = t[t[:, 0] != 0].double()
print("CTAs", t.shape[0], "kernel ms", e0.elapsed_time(e1))
names = ["start->first data", "first data->last MMA issued", "last MMA issued->acc0 ready", "acc0 epilogue", "acc1 ready after acc0 ready",
         "acc1 epilogue", "whole CTA"]
vals = [t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2], t[:, 4] - t[:, 3], t[:, 5] - t[:, 3], t[:, 6] - t[:, 5], t[:, 7] - t[:, 0]]
for n, v in zip(names, vals):
    print(f"{n:36s} mean {v.mean():10.0f}  p10 {v.quantile(0.1):10.0f}  p90 {v.quantile(0.9):10.0f} cycles")
iters = 27 * C // 64
sp = int(os.environ.get("SPLITS", "1"))
iters = iters // sp
print("iters", iters, "ideal MMA cycles per CTA (MT=2,BN=256):", iters * 8 * 128, " mainloop/ideal:", ((t[:, 2] - t[:, 1]).mean() / (iters * 8 * 128)).item())
