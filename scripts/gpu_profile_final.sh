#!/bin/bash
# ncu evidence for the round: launch list of the bench command + --set full captures of the three hot kernels
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 $B > gpurun_out/bench_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/bench_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_list.log 2>&1; echo "list $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3d_tc_kernel -s 2 -c 2 -o gpurun_out/prof_conv $B > gpurun_out/ncu_conv.log 2>&1; echo "conv $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc2 -s 0 -c 1 -o gpurun_out/prof_attn $B > gpurun_out/ncu_attn.log 2>&1; echo "attn $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 3 -c 1 -o gpurun_out/prof_gn $B > gpurun_out/ncu_gn.log 2>&1; echo "gn $?"
ls -la gpurun_out/*.ncu-rep
