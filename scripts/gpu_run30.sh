#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --timeout 600 -m gpu -p no:cacheprovider"
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" >> gpurun_out/$name.log; tail -n 8 gpurun_out/$name.log | cut -c1-300; return $rc; }
run t_kernels $PT tests/test_gpu_kernels.py
run t_models $PT tests/test_gpu_models.py
B200LDM_CONV_AUTOTUNE_LOG=1 TOPN=45 run prof32_auto python scripts/profile_models.py 32g 20 8
B200LDM_CONV_AUTOTUNE_LOG=1 TOPN=12 run prof16_auto python scripts/profile_models.py 16g 36,44,28 8
run sample python bench.py --workload sample_32g --T 200 --steps 2 --warmup 1
grep -h "autotune" gpurun_out/prof16_auto.log gpurun_out/prof32_auto.log | sort | uniq > gpurun_out/autotune_choices.log
