"""One eager AutoencoderKL reconstruct between cudaProfilerStart/Stop, for ncu captures of its kernels:
  ncu --profile-from-start off --set full --clock-control none -k regex:attention_tc_kernel -c 1 -o out python scripts/ncu_ae.py [edge]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-latent-diffusion-model_b200"))
import torch
import b200ldm
import bench as B

edge = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ae = B.seeded_init(b200ldm.AutoencoderKL(**B.AE_32G, precision="bf16"), 4321).to("cuda").eval()
v = torch.rand((1, 1, edge, edge, edge), device="cuda")
for _ in range(2):
    ae.reconstruct(v)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ae.reconstruct(v)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one reconstruct")
