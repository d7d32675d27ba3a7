#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --timeout 600 -m gpu -p no:cacheprovider"
run() { name=$1; shift; echo "=== $name" ; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" >> gpurun_out/$name.log; tail -n 6 gpurun_out/$name.log | cut -c1-900; return $rc; }
run t_kernels $PT tests/test_gpu_kernels.py
run attn_bench python scripts/attn_bench.py
cat gpurun_out/attn_bench.log
run t_models  $PT tests/test_gpu_models.py -s
run bench     python bench.py --steps 10 --warmup 3 --no-cpu-baseline
B200LDM_ATTN_F16V=0 run bench_nof16v python bench.py --steps 10 --warmup 3 --no-cpu-baseline
