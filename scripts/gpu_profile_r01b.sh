#!/bin/bash
# ncu evidence (second pass of round 1): launch list of the bench command + --set full captures of the hot kernels.
# The plain run measures the conv kernel choices and stores them; the profiler runs preload them (no tuning under ncu).
mkdir -p gpurun_out
export B200LDM_CONV_TUNE_FILE=$PWD/gpurun_out/conv_tune.txt
rm -f $B200LDM_CONV_TUNE_FILE
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sampling"
timeout 900 $B > gpurun_out/bench_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/bench_plain.log; exit 1; }
wc -l $B200LDM_CONV_TUNE_FILE
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_list.log 2>&1; echo "list $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3d_tc_kernel -s 2 -c 2 -o gpurun_out/prof_conv $B > gpurun_out/ncu_conv.log 2>&1; echo "conv $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3d_tcp_kernel -s 3 -c 3 -o gpurun_out/prof_convp $B > gpurun_out/ncu_convp.log 2>&1; echo "convp $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc2 -s 0 -c 1 -o gpurun_out/prof_attn $B > gpurun_out/ncu_attn.log 2>&1; echo "attn $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 3 -c 1 -o gpurun_out/prof_gn $B > gpurun_out/ncu_gn.log 2>&1; echo "gn $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tap_gather -s 0 -c 1 -o gpurun_out/prof_tap $B > gpurun_out/ncu_tap.log 2>&1; echo "tap $?"
ls -la gpurun_out/*.ncu-rep
