#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --timeout 600 -m gpu -p no:cacheprovider"
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" >> gpurun_out/$name.log; tail -n 8 gpurun_out/$name.log | cut -c1-300; return $rc; }
run t_kernels $PT tests/test_gpu_kernels.py
run t_models $PT tests/test_gpu_models.py -s
: > gpurun_out/ab_tap.log
for i in 1 2; do
  for arm in "B200LDM_TAP_EXPAND=0" "B200LDM_TAP_EXPAND=1"; do
    env $arm python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sampling 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$arm', 'ms/step %.3f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'], 'conv %.3f' % d['kernel_breakdown']['conv3d']['ms_per_step'], 'clk', d['clocks']['sm_mhz'])
" | tee -a gpurun_out/ab_tap.log
  done
done
TOPN=12 run prof32 python scripts/profile_models.py 32g 20 8
run bench_full python bench.py --steps 20 --warmup 3
