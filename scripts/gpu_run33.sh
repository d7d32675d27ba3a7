#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --timeout 600 -m gpu -p no:cacheprovider"
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" >> gpurun_out/$name.log; tail -n 12 gpurun_out/$name.log | cut -c1-300; return $rc; }
run t_kernels $PT tests/test_gpu_kernels.py
run attn_bench python scripts/attn_bench.py
run t_models $PT tests/test_gpu_models.py
run bench python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sampling
