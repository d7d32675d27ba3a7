#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --timeout 600 -m gpu -p no:cacheprovider"
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" >> gpurun_out/$name.log; tail -n 8 gpurun_out/$name.log | cut -c1-300; return $rc; }
run t_kernels $PT tests/test_gpu_kernels.py
run t_models $PT tests/test_gpu_models.py
for i in 1 2 3; do python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sampling 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
kb = d['kernel_breakdown']
print('ms/step %.3f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'], 'conv %.3f' % kb['conv3d']['ms_per_step'], 'gn_apply %.3f (%.0f GB/s)' % (kb['gn_apply']['ms_per_step'], kb['gn_apply']['GBps']), 'attn %.3f' % kb['attention']['ms_per_step'])
"; done | tee gpurun_out/bench3.log
