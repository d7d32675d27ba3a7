#!/bin/bash
# refresh of the launch list + the attention capture after the default head-dim-64 kernel changed
mkdir -p gpurun_out
export B200LDM_CONV_TUNE_FILE=$PWD/gpurun_out/conv_tune.txt
rm -f $B200LDM_CONV_TUNE_FILE
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sampling"
timeout 900 $B > gpurun_out/bench_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/bench_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_list.log 2>&1; echo "list $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 0 -c 1 -o gpurun_out/prof_attn64 $B > gpurun_out/ncu_attn64.log 2>&1; echo "attn $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 3 -c 1 -o gpurun_out/prof_gn2 $B > gpurun_out/ncu_gn2.log 2>&1; echo "gn $?"
ls -la gpurun_out/prof_attn64.ncu-rep gpurun_out/prof_gn2.ncu-rep
