#!/bin/bash
# round 2, call V: stream-K as a measured candidate (4 % margin): 32g and 16g steps with / without
mkdir -p gpurun_out
for sk in 0 1; do
  echo "=== 32g step, STREAMK=$sk streams=2"
  B200LDM_CONV_STREAMK=$sk B200LDM_CONV_AUTOTUNE_LOG=1 PROFILE_STREAMS=2 TOPN=30 timeout 600 python scripts/profile_models.py 32g 20 8 > gpurun_out/v_32g_sk$sk.log 2> gpurun_out/v_32g_sk$sk.err
  grep -E "graph replay|profiled kernels" gpurun_out/v_32g_sk$sk.log | head -2
  grep -c "kernel 5" gpurun_out/v_32g_sk$sk.err
  echo "=== 16g step, STREAMK=$sk streams=2"
  B200LDM_CONV_STREAMK=$sk B200LDM_CONV_AUTOTUNE_LOG=1 PROFILE_STREAMS=2 TOPN=30 timeout 600 python scripts/profile_models.py 16g 36,44,28 8 > gpurun_out/v_16g_sk$sk.log 2> gpurun_out/v_16g_sk$sk.err
  grep -E "graph replay|profiled kernels" gpurun_out/v_16g_sk$sk.log | head -2
  grep "kernel 5" gpurun_out/v_16g_sk$sk.err | cut -c18-200
done
