#!/bin/bash
# round 2, call O: split-K slabs through smem + TMA stores (B200LDM_CONV_SLAB_TMA), narrower Cout tiles as a measured choice
mkdir -p gpurun_out
echo "=== conv kernel tests"
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "conv or split or gn or norm" 2>&1 | tail -4
for v in 0 1; do
  echo "=== 32g step, SLAB_TMA=$v"
  B200LDM_CONV_SLAB_TMA=$v B200LDM_CONV_AUTOTUNE_LOG=1 PROFILE_STREAMS=2 TOPN=30 timeout 600 python scripts/profile_models.py 32g 20 8 > gpurun_out/o_32g_slab$v.log 2> gpurun_out/o_32g_slab$v.err
  grep -E "graph replay|profiled kernels|decode eager" gpurun_out/o_32g_slab$v.log
  grep -c "Cout tile" gpurun_out/o_32g_slab$v.err
done
echo "=== model tests"
timeout 900 python -m pytest tests/test_gpu_models.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
