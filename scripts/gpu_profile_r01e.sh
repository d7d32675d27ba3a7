#!/bin/bash
# end-of-round refresh: ncu launch list of the default bench run (top-level sub-batch streams on) + one --set full
# capture of the dominant conv kernel
mkdir -p gpurun_out
export B200LDM_CONV_TUNE_FILE=$PWD/gpurun_out/conv_tune_e.txt
rm -f $B200LDM_CONV_TUNE_FILE
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sampling"
timeout 900 $B > gpurun_out/bench_plain_e.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/bench_plain_e.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_e.csv $B > gpurun_out/ncu_list_e.log 2>&1; echo "list $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3d_tc_kernel -s 40 -c 1 -o gpurun_out/prof_conv_e $B > gpurun_out/ncu_conv_e.log 2>&1; echo "conv $?"
ls -la gpurun_out/prof_conv_e.ncu-rep
