#!/bin/bash
# round 2, call J: Q in TMEM on top of P in TMEM
mkdir -p gpurun_out
echo "=== tests QTMEM=1"
B200LDM_ATTN_QTMEM=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "attention" 2>&1 | tail -5
for q in 1 0 1 0; do
  echo "=== QTMEM=$q"
  B200LDM_ATTN_QTMEM=$q timeout 300 python scripts/attn_bench.py 2>&1 | head -5 | tee gpurun_out/attn_qt$q.log
done
