#!/bin/bash
# generic A/B of two env settings on the headline bench: bash scripts/gpu_ab2.sh "ENV_A=.." "ENV_B=.." [reps]
mkdir -p gpurun_out
A="$1"; B="$2"; R=${3:-3}
: > gpurun_out/ab2.log
for i in $(seq 1 $R); do
  for arm in "$A" "$B"; do
    env $arm python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sampling 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$arm', 'ms/step %.3f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'], 'conv %.3f' % d['kernel_breakdown']['conv3d']['ms_per_step'], 'clk', d['clocks']['sm_mhz'])
" | tee -a gpurun_out/ab2.log
  done
done
