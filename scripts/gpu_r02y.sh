#!/bin/bash
# round 2, call Y: split-KV attention - tests, AutoencoderKL reconstruct sweep with / without
mkdir -p gpurun_out
echo "=== attention tests"
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "attention" 2>&1 | tail -5
for rep in 1 2; do
for sp in 0 -1; do
  echo "=== AE reconstruct, ATTN_SPLIT=$sp (rep $rep)"
  for e in 96 128 160 192; do
    B200LDM_ATTN_SPLIT=$sp timeout 300 python scripts/ae_profile_recon.py $e 2>&1 | grep -E "attention  B|reconstruct eager"
  done
done
done
