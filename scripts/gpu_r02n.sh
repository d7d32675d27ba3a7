#!/bin/bash
# round 2, call N: rest of the GPU suite after the tolerance fix; P / Q in TMEM for the wide heads (d = 128 / 256)
mkdir -p gpurun_out
echo "=== full gpu suite"
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
for w in 0 1 2; do
  echo "=== WIDE_TMEM=$w tests"
  B200LDM_ATTN_WIDE_TMEM=$w timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "attention" 2>&1 | tail -2
  echo "=== WIDE_TMEM=$w bench"
  B200LDM_ATTN_WIDE_TMEM=$w timeout 300 python scripts/attn_bench.py 2>&1 | tail -2
done
