"""Where the softmax warpgroups of attention_tc2_kernel spend their cycles (clock64 counters per CTA / warpgroup)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
import torch
from b200ldm import _lib, ops

lib = _lib.load()
lib.b200ldm_debug_set_conv_trace.argtypes = [ctypes.c_void_p]
dev = "cuda"
B, H, L, d = (int(v) for v in os.environ.get("SHAPE", "8,4,5544,64").split(","))
qkv = torch.randn((B, L, 3 * H * d), device=dev).to(torch.bfloat16)
for _ in range(3):
    ops.attention(qkv, H, d, d ** -0.5)
torch.cuda.synchronize()
trace = torch.zeros((2048, 8), dtype=torch.int64, device=dev)   # per CTA: 8 softmax-warp slots, then 8 MMA-thread slots
lib.b200ldm_debug_set_conv_trace(trace.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.attention(qkv, H, d, d ** -0.5); e1.record()
torch.cuda.synchronize()
lib.b200ldm_debug_set_conv_trace(None)
tall = trace.cpu().double().reshape(1024, 16)
tall = tall[tall[:, 6] != 0]
t, tm = tall[:, :8], tall[:, 8:]
print(f"B{B} H{H} L{L} d{d}: {e0.elapsed_time(e1):.3f} ms, traced warpgroups {t.shape[0]}, dbg={os.environ.get('B200LDM_ATTN_DBG', '0')}")
names = ["wait s_full", "TMEM load S", "max", "wait pv_done", "rescale", "exp + P store"]
per_tile = t[:, :6] / t[:, 6:7]
for i, n in enumerate(names):
    print(f"   {n:16s} {per_tile[:, i].mean():8.0f} cycles/tile (p10 {per_tile[:, i].quantile(0.1):.0f}, p90 {per_tile[:, i].quantile(0.9):.0f})")
print(f"   total/tile       {(t[:, 7] / t[:, 6]).mean():8.0f}")
if tm[:, 6].sum() > 0:
    print("   MMA-issuing thread, cycles/tile:")
    for i, n in enumerate(["wait K_full", "wait S buffer free", "issue S (MMAs + 2 commits)", "wait V_full", "wait P_full", "issue PV (MMAs + 2 commits)"]):
        v = tm[:, i] / tm[:, 6]
        print(f"      {n:28s} {v.mean():8.0f} (p10 {v.quantile(0.1):.0f}, p90 {v.quantile(0.9):.0f})")
    print(f"      sum                          {(tm[:, :6].sum(1) / tm[:, 6]).mean():8.0f}")
