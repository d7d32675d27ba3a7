#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --timeout 600 -m gpu -p no:cacheprovider"
run() { name=$1; shift; echo "=== $name" ; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" >> gpurun_out/$name.log; tail -n 8 gpurun_out/$name.log | cut -c1-600; return $rc; }
run t_gn $PT tests/test_gpu_kernels.py -k "groupnorm"
run t_models  $PT tests/test_gpu_models.py -s
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
if run bench_plain $B; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 3 -c 2 -o gpurun_out/prof_gn $B > gpurun_out/ncu_gn.log 2>&1; echo "ncu gn exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 0 -c 1 -o gpurun_out/prof_attn $B > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3d_tc_kernel -s 1 -c 1 -o gpurun_out/prof_conv2 $B > gpurun_out/ncu_conv2.log 2>&1; echo "ncu conv exit $?"
fi
