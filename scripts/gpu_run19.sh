#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --timeout 600 -m gpu -p no:cacheprovider"
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" >> gpurun_out/$name.log; tail -n 30 gpurun_out/$name.log | cut -c1-400; return $rc; }
run t_models  $PT tests/test_gpu_models.py -s
run t_sched   $PT tests/test_gpu_kernels.py -k "sched or philox"
run sample_breakdown python scripts/sample_breakdown.py 200 20
run sample python bench.py --workload sample_32g --T 200 --steps 2 --warmup 1
