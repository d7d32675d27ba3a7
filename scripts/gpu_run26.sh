#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --timeout 600 -m gpu -p no:cacheprovider"
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" >> gpurun_out/$name.log; tail -n 12 gpurun_out/$name.log | cut -c1-300; return $rc; }
B200LDM_CONV_KERNEL=2 run t_conv_tcp $PT tests/test_gpu_kernels.py -k "conv3d or linear" -x
B200LDM_CONV_KERNEL=2 run t_models_tcp $PT tests/test_gpu_models.py -x
B200LDM_CONV_KERNEL=1 run t_conv_tc2 $PT tests/test_gpu_kernels.py -k "conv3d or linear" -x
B200LDM_CONV_AUTOTUNE_LOG=1 TOPN=40 run prof16_auto python scripts/profile_models.py 16g 36,44,28 8
B200LDM_CONV_AUTOTUNE_LOG=1 TOPN=40 run prof32_auto python scripts/profile_models.py 32g 20 8
run bench_auto python bench.py --steps 10 --warmup 3 --no-cpu-baseline
B200LDM_CONV_KERNEL=0 run bench_k0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
grep -h "autotune" gpurun_out/prof16_auto.log gpurun_out/prof32_auto.log | sort | uniq > gpurun_out/autotune_choices.log
