#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --timeout 600 -m gpu -p no:cacheprovider"
run() { name=$1; shift; echo "=== $name" ; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" >> gpurun_out/$name.log; tail -n 6 gpurun_out/$name.log | cut -c1-900; return $rc; }
run t_kernels $PT tests/test_gpu_kernels.py
B200LDM_CONV_MT=1 B200LDM_CONV_CL=2 run t_conv_mt1cl2 $PT tests/test_gpu_kernels.py -k "conv3d"
run t_models  $PT tests/test_gpu_models.py -s
run bench     python bench.py --steps 10 --warmup 3 --no-cpu-baseline
python scripts/profile_models.py 32g 20 8 2>&1 | grep -E "==|graph|decode"
TOPN=22 python scripts/profile_models.py 16g 36,44,28 8 2>&1 | grep -E "conv3d " | head -24
