#!/bin/bash
# round 2, call X: is the replayed step power-bound?  stream-K on / off, alternating, ms/step next to SM clock and board power
mkdir -p gpurun_out
for rep in 1 2; do
  for sk in 0 1; do
    B200LDM_CONV_STREAMK=$sk timeout 600 python scripts/power_ab.py 32g 600 2>&1 | grep "replays"
  done
done
for sk in 0 1; do
  B200LDM_CONV_STREAMK=$sk timeout 600 python scripts/power_ab.py 16g 200 2>&1 | grep "replays"
done
