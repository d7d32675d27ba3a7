#!/bin/bash
# round 2, call AC (final evidence run): full GPU suite + smoke + default bench + reference arm; ncu launch lists of the
# 16g / 32g steps; ncu --set full of the dominant conv launch (roofline.traffic), the attention kernel, GroupNorm apply
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err; rc=$?; echo "exit $rc" >> gpurun_out/$name.err; tail -n 4 gpurun_out/$name.log | cut -c1-400; tail -n 2 gpurun_out/$name.err | cut -c1-300; return $rc; }
TMO=1500 run ac_gpu python -m pytest tests -x -q -m gpu -p no:cacheprovider
TMO=300 run ac_smoke python __graft_entry__.py smoke
TMO=900 run ac_bench python bench.py
python scripts/show_bench.py gpurun_out/ac_bench.log 2>&1 | head -30
TMO=600 run ac_bench_ref python bench.py --impl reference --steps 2 --warmup 1
export B200LDM_CONV_TUNE_FILE=$PWD/gpurun_out/conv_tune_r02ac.txt
rm -f $B200LDM_CONV_TUNE_FILE
TMO=300 run ac_step16_plain python scripts/ncu_step.py 16g 36,44,28 8
TMO=300 run ac_step32_plain python scripts/ncu_step.py 32g 20 8
TMO=600 run ac_step16_ncu ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02ac_launches_16g_step.csv python scripts/ncu_step.py 16g 36,44,28 8
TMO=600 run ac_step32_ncu ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02ac_launches_32g_step.csv python scripts/ncu_step.py 32g 20 8
TMO=900 run ac_conv_full ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv3d_tc_kernel -s 2 -c 1 -f -o gpurun_out/r02ac_conv_full python scripts/ncu_step.py 16g 36,44,28 8
TMO=900 run ac_attn_full ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 1 -c 1 -f -o gpurun_out/r02ac_attn_full python scripts/ncu_step.py 16g 36,44,28 8
TMO=900 run ac_gn_full ncu --profile-from-start off --set full --clock-control none -k regex:gn_apply_kernel -s 2 -c 1 -f -o gpurun_out/r02ac_gn_full python scripts/ncu_step.py 16g 36,44,28 8
ls -la gpurun_out/r02ac_*.ncu-rep
