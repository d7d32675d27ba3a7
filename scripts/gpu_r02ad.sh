#!/bin/bash
# round 2, call AD: every tuner candidate forced in turn
mkdir -p gpurun_out
for c in "2 1,1,24 128 0 384 1 1" "3 5,3,7 256 0 512 3 1" "8 5,5,5 256 0 512 3 1" "1 9,11,7 256 0 256 1 1" "3 10,10,10 128 0 256 3 1"; do
  echo "=== $c"
  timeout 120 python scripts/conv_candidates.py $c 2>&1 | awk '{ if ($4+0 > 0.01 || $0 !~ /relerr/) print }' | tail -8
done
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "every_tuner_candidate or split_k_ or stream_k" 2>&1 | tail -8
