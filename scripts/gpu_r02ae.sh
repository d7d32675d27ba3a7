#!/bin/bash
# round 2, call AE: head-dim-256 attention with two MMA-issuing threads - AutoencoderKL reconstruct with / without
mkdir -p gpurun_out
for rep in 1 2; do
for t in 0 1; do
  for e in 128 160; do
    echo "=== TWO_ISSUERS=$t edge $e (rep $rep)"
    B200LDM_ATTN_TWO_ISSUERS=$t timeout 300 python scripts/ae_profile_recon.py $e > gpurun_out/ae_ti${t}_$e.log 2>&1
    grep -E "attention  B|reconstruct eager" gpurun_out/ae_ti${t}_$e.log
  done
done
done
timeout 600 python -m pytest tests/test_gpu_long.py tests/test_gpu_models.py -x -q -m gpu -p no:cacheprovider -k "cfg5 or autoencoder or multigpu or optimized or stable" 2>&1 | tail -3
