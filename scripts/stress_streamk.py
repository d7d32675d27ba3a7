"""Forward-progress check of the stream-K conv kernel under concurrency: 2 / 4 streams launch stream-K convs (148 CTAs each, owners
polling their contributors' flags) at the same time, repeatedly; results must equal the single-stream result bit for bit.
python scripts/stress_streamk.py [iterations]"""
import ctypes, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
import torch
from b200ldm import _lib, ops

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
lib = _lib.load()
lib.b200ldm_debug_set_conv_kernel.argtypes = [ctypes.c_int]
dev = "cuda"
shapes = [(8, (10, 10, 10), 256, 256), (4, (20, 20, 20), 256, 256), (8, (18, 22, 14), 256, 256), (8, (5, 5, 5), 512, 512)]
probs = []
for N, sp, C, Cout in shapes:
    x = torch.randn((N,) + sp + (C,), device=dev).bfloat16()
    w = ops.repack_conv_weight(torch.randn((Cout, C, 3, 3, 3), device=dev) / math.sqrt(C * 27), Cout, C, torch.bfloat16)
    b = torch.zeros(Cout, device=dev)
    probs.append((x, w, b, Cout))
lib.b200ldm_debug_set_conv_kernel(5)
try:
    refs = [ops.conv3d(x, w, b, co) for x, w, b, co in probs]
    torch.cuda.synchronize()
    for nstreams in (2, 4):
        streams = [torch.cuda.Stream() for _ in range(nstreams)]
        bad = 0
        for it in range(iters):
            outs = []
            for si, st in enumerate(streams):
                with torch.cuda.stream(st):
                    k = (it + si) % len(probs)
                    x, w, b, co = probs[k]
                    outs.append((k, ops.conv3d(x, w, b, co)))
            torch.cuda.synchronize()
            bad += sum(0 if torch.equal(y, refs[k]) else 1 for k, y in outs)
        print(f"{nstreams} concurrent streams x {iters} rounds of stream-K convs: {bad} mismatches")
        assert bad == 0
finally:
    lib.b200ldm_debug_set_conv_kernel(-1)
print("ok")
