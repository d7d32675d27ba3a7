#!/bin/bash
# tests + first bench line + ncu launch list + one ncu --set full capture of the dominant conv kernel
mkdir -p gpurun_out
PT="python -m pytest -q --timeout 600 -m gpu -p no:cacheprovider"
run() { name=$1; shift; echo "=== $name" ; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" >> gpurun_out/$name.log; tail -n 30 gpurun_out/$name.log; return $rc; }
run t_att_tc  $PT tests/test_gpu_kernels.py -k "attention_tcgen05"
run t_kernels $PT tests/test_gpu_kernels.py -k "not attention_tcgen05"
run t_models  $PT tests/test_gpu_models.py -s
run bench     python bench.py --steps 10 --warmup 3
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
if run bench_plain $B; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_list.log 2>&1
  echo "ncu list exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3d_tc_kernel -s 1 -c 3 -o gpurun_out/prof_conv $B > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
ls -la gpurun_out | head -30
