#!/bin/bash
# round 2, call S: default bench (cfg1 / configs[3] / configs[4] legs) + reference arm + ncu launch list of one 32g step
mkdir -p gpurun_out
echo "=== bench"
t0=$(date +%s)
timeout 1200 python bench.py > gpurun_out/s_bench.log 2> gpurun_out/s_bench.err; echo "exit $? after $(( $(date +%s) - t0 )) s"
tail -2 gpurun_out/s_bench.err
python scripts/show_bench.py gpurun_out/s_bench.log 2>&1 | head -40
echo "=== reference arm"
t0=$(date +%s)
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s_bench_ref.log 2> gpurun_out/s_bench_ref.err; echo "exit $? after $(( $(date +%s) - t0 )) s"
cut -c1-400 gpurun_out/s_bench_ref.log
echo "=== 32g launch list"
export B200LDM_CONV_TUNE_FILE=$PWD/gpurun_out/conv_tune_r02s.txt
rm -f $B200LDM_CONV_TUNE_FILE
timeout 300 python scripts/ncu_step.py 32g 20 8 > gpurun_out/s_step32_plain.log 2>&1; echo "plain exit $?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02s_launches_32g_step.csv python scripts/ncu_step.py 32g 20 8 > gpurun_out/s_step32_ncu.log 2>&1; echo "ncu exit $?"
python scripts/launch_summary.py gpurun_out/r02s_launches_32g_step.csv 2>&1 | head -24
