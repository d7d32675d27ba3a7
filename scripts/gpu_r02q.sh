#!/bin/bash
# round 2, call Q: clock64 timeline of the split-K convs of the small levels
mkdir -p gpurun_out
for cfg in "5,5,5 1024 1024 9" "10,10,10 512 512 2"; do
  set -- $cfg
  echo "=== trace DHW=$1 $2->$3"
  DHW=$1 CIN=$2 COUT=$3 SPLITS=$4 B200LDM_CONV_AUTOTUNE_LOG=1 timeout 300 python scripts/conv_trace.py 2>&1 | tail -12
done
