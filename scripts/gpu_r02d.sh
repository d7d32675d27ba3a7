#!/bin/bash
# round 2, call D: clock64 phase trace of the softmax warps of the head-dim-64 attention kernel (traced build)
mkdir -p gpurun_out
export B200LDM_NVCC_EXTRA=-DB200LDM_ATTN_TRACE
timeout 900 python 3d-latent-diffusion-model_b200/build.py --force > gpurun_out/build_trace.log 2>&1 || { tail -5 gpurun_out/build_trace.log; exit 1; }
unset B200LDM_NVCC_EXTRA
for m in 1 0 4 5; do
  echo "=== trace mode $m"
  B200LDM_ATTN_PACKED=$m timeout 300 python scripts/attn_trace.py 2>&1 | tee gpurun_out/attn_trace_mode$m.log
done
echo "=== trace mode 1, exp2 replaced by a multiply (dbg=1)"
B200LDM_ATTN_DBG=1 B200LDM_ATTN_PACKED=1 timeout 300 python scripts/attn_trace.py 2>&1 | tee gpurun_out/attn_trace_mode1_nomufu.log
echo "=== trace mode 1 L1000"
SHAPE=8,8,1000,64 B200LDM_ATTN_PACKED=1 timeout 300 python scripts/attn_trace.py 2>&1 | tee gpurun_out/attn_trace_mode1_L1000.log
