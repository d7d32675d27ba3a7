#!/bin/bash
# round 2, call Z: energy experiment - sustained 16g step with the CTA-pair kernel / 2-CTA weight multicast pinned
mkdir -p gpurun_out
for rep in 1 2; do
  echo "--- default";            timeout 600 python scripts/power_ab.py 16g 200 2>&1 | grep "200 replays"
  echo "--- CONV_CL=2";          B200LDM_CONV_CL=2 timeout 600 python scripts/power_ab.py 16g 200 2>&1 | grep "200 replays"
  echo "--- CONV_KERNEL=1 pairs"; B200LDM_CONV_KERNEL=1 timeout 600 python scripts/power_ab.py 16g 200 2>&1 | grep "200 replays"
done
