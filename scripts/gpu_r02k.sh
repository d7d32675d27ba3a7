#!/bin/bash
# round 2, call K: full GPU test suite + smoke + bench with the round's defaults; ncu --set full of the dominant conv
# launch (roofline.traffic) and of the attention kernel; launch list of the default bench step
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err; rc=$?; echo "exit $rc" >> gpurun_out/$name.err; tail -n 4 gpurun_out/$name.log | cut -c1-500; tail -n 3 gpurun_out/$name.err | cut -c1-300; return $rc; }
TMO=1500 run tk_gpu python -m pytest tests -x -q -m gpu -p no:cacheprovider
TMO=300 run smoke_k python __graft_entry__.py smoke
TMO=900 run bench_k python bench.py
TMO=600 run bench_ref_k python bench.py --impl reference --steps 2 --warmup 1
export B200LDM_CONV_TUNE_FILE=$PWD/gpurun_out/conv_tune_r02k.txt
rm -f $B200LDM_CONV_TUNE_FILE
TMO=300 run step16k_plain python scripts/ncu_step.py 16g 36,44,28 8
TMO=600 run step16k_ncu ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02k_launches_16g_step.csv python scripts/ncu_step.py 16g 36,44,28 8
TMO=900 run conv_full_ncu ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv3d_tc_kernel -s 2 -c 1 -f -o gpurun_out/r02k_conv_full python scripts/ncu_step.py 16g 36,44,28 8
TMO=900 run attn_full_ncu ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 1 -c 1 -f -o gpurun_out/r02k_attn_full python scripts/ncu_step.py 16g 36,44,28 8
TMO=900 run gn_full_ncu ncu --profile-from-start off --set full --clock-control none -k regex:gn_apply_kernel -s 2 -c 1 -f -o gpurun_out/r02k_gn_full python scripts/ncu_step.py 16g 36,44,28 8
ls -la gpurun_out/r02k_*.ncu-rep
