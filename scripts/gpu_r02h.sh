#!/bin/bash
# round 2, call H: single S buffer + P in TMEM, three CTAs per SM
mkdir -p gpurun_out
echo "=== tests SB=1"
B200LDM_ATTN_SB=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "attention" 2>&1 | tail -5
for cfg in "1 1" "1 3" "1 4" "0 1"; do
  set -- $cfg
  echo "=== SB=$1 PACKED=$2 (PTMEM=1)"
  B200LDM_ATTN_SB=$1 B200LDM_ATTN_PTMEM=1 B200LDM_ATTN_PACKED=$2 timeout 300 python scripts/attn_bench.py 2>&1 | head -5 | tee gpurun_out/attn_sb$1_mode$2.log
done
echo "=== SB=1 PACKED=1 stagger 500"
B200LDM_ATTN_SB=1 B200LDM_ATTN_STAGGER=500 timeout 300 python scripts/attn_bench.py 2>&1 | head -5
