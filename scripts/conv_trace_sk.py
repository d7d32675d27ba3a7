"""Per-CTA timeline of the stream-K conv kernel (conv_sk.cu, clock64 stamps): DHW=20,20,20 CIN=256 COUT=256 python scripts/conv_trace_sk.py"""
import ctypes, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
import torch
from b200ldm import _lib, ops

lib = _lib.load()
lib.b200ldm_debug_set_conv_trace.argtypes = [ctypes.c_void_p]
lib.b200ldm_debug_set_conv_kernel.argtypes = [ctypes.c_int]
dev = "cuda"
D, H, W = (int(v) for v in os.environ.get("DHW", "20,20,20").split(","))
N, C, Cout = 8, int(os.environ.get("CIN", 256)), int(os.environ.get("COUT", 256))
x = torch.randn((N, D, H, W, C), device=dev).bfloat16()
w = ops.repack_conv_weight(torch.randn((Cout, C, 3, 3, 3), device=dev) / math.sqrt(C * 27), Cout, C, torch.bfloat16)
b = torch.zeros(Cout, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for kern in (0, 5):
    lib.b200ldm_debug_set_conv_kernel(kern)
    for _ in range(3):
        y = ops.conv3d(x, w, b, Cout)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        y = ops.conv3d(x, w, b, Cout)
    e1.record(); torch.cuda.synchronize()
    print(f"kernel {kern}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per launch (10 back to back)")
trace = torch.zeros((4096, 16), dtype=torch.int64, device=dev)
lib.b200ldm_debug_set_conv_trace(trace.data_ptr())
y = ops.conv3d(x, w, b, Cout)
torch.cuda.synchronize()
lib.b200ldm_debug_set_conv_trace(None)
lib.b200ldm_debug_set_conv_kernel(-1)
t = trace.cpu()
t = t[t[:, 0] != 0].double()
print("CTAs", t.shape[0])
def show(name, v):
    print(f"{name:44s} mean {v.mean():9.0f}  p10 {v.quantile(0.1):9.0f}  p90 {v.quantile(0.9):9.0f}  max {v.max():9.0f}")
show("start -> first stage landed", t[:, 1] - t[:, 0])
for i in range(4):
    m = t[:, 3 + 3 * i] != 0
    if m.sum() == 0:
        break
    u = t[m]
    print(f"segment {i}: {int(m.sum())} CTAs")
    if i > 0:
        show("  prev last MMA issued -> first MMA issued", u[:, 2 + 3 * i] - u[:, 3 + 3 * (i - 1)])
    show("  first MMA -> last MMA issued", u[:, 3 + 3 * i] - u[:, 2 + 3 * i])
    show("  last MMA issued -> epilogue done", u[:, 4 + 3 * i] - u[:, 3 + 3 * i])
show("whole CTA", t[:, 15] - t[:, 0])
iters = 27 * C // 64
tiles = (N * ((D + 4) // 5) * ((H + 4) // 5) * ((W + 4) // 5) + 1) // 2 * (Cout // 256)
print("ideal MMA cycles per CTA:", tiles * iters * 1024 / t.shape[0])
