#!/bin/bash
# round 2, call AB: tap-packed conv_in - tests (kernel + AutoencoderKL model parity), reconstruct sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py tests/test_gpu_long.py -x -q -m gpu -p no:cacheprovider -k "packed_taps or autoencoder or ae or cfg5 or encode or call" 2>&1 | tail -3
for tp in 0 1; do
  for e in 160; do
    echo "=== TAP_PACK=$tp edge $e"
    B200LDM_TAP_PACK=$tp timeout 300 python scripts/ae_profile_recon.py $e 2>&1 | grep -E "pack_taps|64->64 k3s1 @|64->64 k1s1 @1x(96|160)|reconstruct eager|profiled kernels"
  done
done
