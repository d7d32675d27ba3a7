#!/bin/bash
# round 2, call E: co-resident CTA stagger sweep of the head-dim-64 attention kernel (+ redundant pv_done wait removed)
mkdir -p gpurun_out
for st in 0 300 600 800 1000 1400 0 800; do
  echo "=== stagger $st"
  B200LDM_ATTN_STAGGER=$st timeout 300 python scripts/attn_bench.py 2>&1 | head -5 | tee -a gpurun_out/attn_stagger_$st.log
done
echo "=== stagger 800 + poly 4"
B200LDM_ATTN_STAGGER=800 B200LDM_ATTN_PACKED=4 timeout 300 python scripts/attn_bench.py 2>&1 | head -5 | tee gpurun_out/attn_stagger_800_poly4.log
echo "=== tests with stagger 800"
B200LDM_ATTN_STAGGER=800 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "attention" 2>&1 | tail -3
