#!/bin/bash
# round 2, call AF: ncu --set full of the new kernels - stream-K conv (16g step, 18x22x14 level), head-dim-256 attention with
# two issuing threads + split-KV (AutoencoderKL reconstruct at 128^3), merge kernel
mkdir -p gpurun_out
export B200LDM_CONV_TUNE_FILE=$PWD/gpurun_out/conv_tune_r02af.txt
rm -f $B200LDM_CONV_TUNE_FILE
timeout 300 python scripts/ncu_step.py 16g 36,44,28 8 > gpurun_out/af_plain16.log 2>&1; echo "plain16 exit $?"
timeout 300 python scripts/ncu_ae.py 128 > gpurun_out/af_plain_ae.log 2>&1; echo "plain ae exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv3d_sk_kernel -s 1 -c 1 -f -o gpurun_out/r02af_sk_full python scripts/ncu_step.py 16g 36,44,28 8 > gpurun_out/af_sk.log 2>&1; echo "sk exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 1 -c 1 -f -o gpurun_out/r02af_attn256_full python scripts/ncu_ae.py 128 > gpurun_out/af_attn.log 2>&1; echo "attn exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:attention_merge_kernel -c 1 -f -o gpurun_out/r02af_merge_full python scripts/ncu_ae.py 128 > gpurun_out/af_merge.log 2>&1; echo "merge exit $?"
ls -la gpurun_out/r02af_*.ncu-rep
