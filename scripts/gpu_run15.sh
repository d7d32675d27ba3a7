#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --timeout 600 -m gpu -p no:cacheprovider"
run() { name=$1; shift; echo "=== $name" ; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" >> gpurun_out/$name.log; tail -n 6 gpurun_out/$name.log | cut -c1-900; return $rc; }
run t_kernels $PT tests/test_gpu_kernels.py
B200LDM_CONV_MT=2 B200LDM_CONV_SPLITS=3 run t_conv_mt2_split3 $PT tests/test_gpu_kernels.py -k "conv3d"
run t_models  $PT tests/test_gpu_models.py -s
run bench     python bench.py --steps 10 --warmup 3 --no-cpu-baseline
python scripts/profile_models.py 32g 20 8 2>&1 | grep -E "==|graph|decode"
