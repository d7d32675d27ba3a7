#!/bin/bash
# round 2, call G: ncu --set full (with source counters) of the head-dim-64 attention kernel, P through TMEM and through smem
mkdir -p gpurun_out
for pt in 1 0; do
  B200LDM_ATTN_PTMEM=$pt timeout 300 python scripts/attn_once.py > gpurun_out/attn_once_pt$pt.log 2>&1 || { echo plain run failed; exit 1; }
  B200LDM_ATTN_PTMEM=$pt timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 2 -c 1 -f -o gpurun_out/r02g_attn64_pt$pt python scripts/attn_once.py > gpurun_out/ncu_attn_pt$pt.log 2>&1; echo "ncu pt=$pt rc=$?"
done
ls -la gpurun_out/r02g_attn64_pt*.ncu-rep
