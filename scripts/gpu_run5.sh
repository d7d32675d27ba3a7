#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --timeout 600 -m gpu -p no:cacheprovider"
run() { name=$1; shift; echo "=== $name" ; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" >> gpurun_out/$name.log; tail -n 8 gpurun_out/$name.log | cut -c1-800; return $rc; }
B200LDM_CONV_MT=2 B200LDM_CONV_CL=2 run t_conv_cl2 $PT tests/test_gpu_kernels.py -k "conv3d"
run t_kernels $PT tests/test_gpu_kernels.py
run t_models  $PT tests/test_gpu_models.py -s
run bench     python bench.py --steps 10 --warmup 3 --no-cpu-baseline
B200LDM_CONV_CL=1 run bench_cl1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
