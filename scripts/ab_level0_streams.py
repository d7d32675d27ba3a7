"""Does splitting the batch over two streams hide the partial last wave of the 32g UNet's level-0 convs
(256 -> 256 k3 @ 8 x 20^3: 256 CTAs = 2 waves for 1.7 waves of work)?  Chain of (GroupNorm+SiLU, conv) x 6 on the full
batch on one stream vs the two half batches on two streams, both graph-captured."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
import torch
from b200ldm import ops

dev, dt = "cuda", torch.bfloat16
C = 256
w = ops.repack_conv_weight((torch.randn(C, C, 3, 3, 3) * (C * 27) ** -0.5).to(dev), C, C, dt)
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
x = torch.randn((8, 20, 20, 20, C), device=dev).to(dt)


def chain(h):
    for _ in range(6):
        h = ops.groupnorm(h, gamma, beta, 32, 1e-6, True)
        h = ops.conv3d(h, w, None, C, 3, 1, 1, 1, want_stats=True)
    return h


def full():
    return chain(x)


side = torch.cuda.Stream()


def halves():
    cur = torch.cuda.current_stream()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        b = chain(x[4:])
    a = chain(x[:4])
    cur.wait_stream(side)
    return a, b


for fn in (full, halves):
    fn(); fn()
torch.cuda.synchronize()
graphs = {}
for name, fn in (("full batch, 1 stream", full), ("2 half batches, 2 streams", halves)):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    graphs[name] = g
for rep in range(3):
    for name, g in graphs.items():
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / 10:.3f} ms per 6 x (GroupNorm + conv)", flush=True)
