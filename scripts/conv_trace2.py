"""Where the persistent CTA-pair conv kernel (conv_tc2.cu) spends its cycles: per-CTA wait / busy counters."""
import ctypes, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
import torch
from b200ldm import _lib, ops

lib = _lib.load()
lib.b200ldm_debug_set_conv_trace.argtypes = [ctypes.c_void_p]
dev = "cuda"
for DHW, C, Cout, k in [((36, 44, 28), 256, 256, 3), ((18, 22, 14), 256, 256, 3), ((18, 22, 14), 256, 256, 1),
                        ((20, 20, 20), 256, 256, 3)]:
    D, H, W = DHW
    N = 8
    x = torch.randn((N, D, H, W, C), device=dev).bfloat16()
    w = ops.repack_conv_weight(torch.randn((Cout, C, k, k, k), device=dev) / math.sqrt(C * k ** 3), Cout, C, torch.bfloat16)
    b = torch.zeros(Cout, device=dev)
    pad = 1 if k == 3 else 0
    for _ in range(3):
        y = ops.conv3d(x, w, b, Cout, ksize=k, pad_lo=pad)
    torch.cuda.synchronize()
    trace = torch.zeros((4096, 8), dtype=torch.int64, device=dev)
    lib.b200ldm_debug_set_conv_trace(trace.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y = ops.conv3d(x, w, b, Cout, ksize=k, pad_lo=pad); e1.record()
    torch.cuda.synchronize()
    lib.b200ldm_debug_set_conv_trace(None)
    t = trace.cpu()
    t = t[t[:, 0] != 0].double()
    ms = e0.elapsed_time(e1)
    fl = 2.0 * N * D * H * W * C * Cout * k ** 3
    print(f"== {C}->{Cout} k{k} @8x{D}x{H}x{W}: {ms:.3f} ms, {fl / ms / 1e9:.0f} TFLOP/s, CTAs {t.shape[0]}")
    if t.shape[0] == 0:
        continue
    lead, peer = t[0::2], t[1::2]
    whole = (t[:, 7] - t[:, 0])
    print(f"   whole CTA cycles mean {whole.mean():.0f} (p10 {whole.quantile(0.1):.0f}, p90 {whole.quantile(0.9):.0f}); "
          f"tiles/pair mean {lead[:, 6].mean():.2f}")
    iters = k ** 3 * C // 64
    print(f"   ideal MMA cycles per tile {iters * 512}; per-pair ideal total {lead[:, 6].mean() * iters * 512:.0f}")
    for nm, col, who in [("MMA wait full", 1, lead), ("MMA wait tmem_empty", 2, lead), ("producer wait empty (leader)", 3, lead),
                         ("producer wait empty (peer)", 3, peer), ("epilogue wait tmem_full", 4, t), ("epilogue busy", 5, t)]:
        v = who[:, col]
        print(f"   {nm:32s} mean {v.mean():10.0f}  p10 {v.quantile(0.1):10.0f}  p90 {v.quantile(0.9):10.0f}")
