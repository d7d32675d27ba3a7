#!/bin/bash
# A/B of two env settings on the headline bench, alternating so slow drifts of the power state hit both arms
mkdir -p gpurun_out
: > gpurun_out/ab.log
for i in 1 2 3; do
  for arm in "B200LDM_CONV_KERNEL=0" "B200LDM_CONV_AUTOTUNE=1"; do
    env $arm python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sampling 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$arm', 'ms/step %.3f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'], 'conv %.3f' % d['kernel_breakdown']['conv3d']['ms_per_step'], 'clk', d['clocks']['sm_mhz'])
" | tee -a gpurun_out/ab.log
  done
done
