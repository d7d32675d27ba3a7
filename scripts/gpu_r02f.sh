#!/bin/bash
# round 2, call F: P through TMEM (TS-mode PV MMA) in the one-query-tile attention kernel
mkdir -p gpurun_out
echo "=== tests PTMEM=1"
B200LDM_ATTN_PTMEM=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "attention" 2>&1 | tail -5
for cfg in "0 1" "1 1" "1 3" "1 4" "1 5" "0 1" "1 1"; do
  set -- $cfg
  echo "=== PTMEM=$1 PACKED=$2"
  B200LDM_ATTN_PTMEM=$1 B200LDM_ATTN_PACKED=$2 timeout 300 python scripts/attn_bench.py 2>&1 | tee gpurun_out/attn_pt$1_mode$2.log
done
echo "=== PTMEM=1 PACKED=4 stagger 800"
B200LDM_ATTN_PTMEM=1 B200LDM_ATTN_PACKED=4 B200LDM_ATTN_STAGGER=800 timeout 300 python scripts/attn_bench.py 2>&1 | head -4
