#!/bin/bash
# round 2, call T: stream-K conv kernel - tests, then the 32g / 16g steps with it among the measured candidates
mkdir -p gpurun_out
echo "=== stream-K tests"
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "stream_k" 2>&1 | tail -6
for sk in 0 1; do
  echo "=== 32g step, STREAMK=$sk streams=2"
  B200LDM_CONV_STREAMK=$sk B200LDM_CONV_AUTOTUNE_LOG=1 PROFILE_STREAMS=2 TOPN=30 timeout 600 python scripts/profile_models.py 32g 20 8 > gpurun_out/t_32g_sk$sk.log 2> gpurun_out/t_32g_sk$sk.err
  grep -E "graph replay|profiled kernels" gpurun_out/t_32g_sk$sk.log | head -3
  grep -c "kernel 5" gpurun_out/t_32g_sk$sk.err
done
echo "=== 32g step, STREAMK=1 streams=1"
B200LDM_CONV_STREAMK=1 PROFILE_STREAMS=1 TOPN=30 timeout 600 python scripts/profile_models.py 32g 20 8 > gpurun_out/t_32g_sk1_s1.log 2> gpurun_out/t_32g_sk1_s1.err
grep -E "graph replay|profiled kernels" gpurun_out/t_32g_sk1_s1.log | head -3
