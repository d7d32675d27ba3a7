#!/bin/bash
# round 2, call I: one CTA per SM (smem padding) vs two: does a second CTA add throughput at all?
mkdir -p gpurun_out
for cfg in "0 0" "0 70000" "1 0" "1 70000"; do
  set -- $cfg
  echo "=== PTMEM=$1 SMEM_PAD=$2"
  B200LDM_ATTN_PTMEM=$1 B200LDM_ATTN_SMEM_PAD=$2 timeout 300 python scripts/attn_bench.py 2>&1 | head -5
done
echo "=== SB=1 with pad 30000 (2 CTAs/SM) and 100000 (1 CTA/SM)"
B200LDM_ATTN_SB=1 B200LDM_ATTN_SMEM_PAD=30000 timeout 300 python scripts/attn_bench.py 2>&1 | head -5
B200LDM_ATTN_SB=1 B200LDM_ATTN_SMEM_PAD=100000 timeout 300 python scripts/attn_bench.py 2>&1 | head -5
