#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --timeout 600 -m gpu -p no:cacheprovider"
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" >> gpurun_out/$name.log; tail -n 25 gpurun_out/$name.log | cut -c1-300; return $rc; }
B200LDM_CONV_TC2=1 run t_conv_tc2 $PT tests/test_gpu_kernels.py -k "conv3d or linear" -x
B200LDM_CONV_TC2=1 run t_models_tc2 $PT tests/test_gpu_models.py -x
B200LDM_CONV_TC2=1 TOPN=30 run prof16_tc2 python scripts/profile_models.py 16g 36,44,28 8
B200LDM_CONV_TC2=0 run bench_tc2_0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
B200LDM_CONV_TC2=1 run bench_tc2_1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run bench_tc2_auto python bench.py --steps 10 --warmup 3 --no-cpu-baseline
