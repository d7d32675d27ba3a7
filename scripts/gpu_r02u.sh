#!/bin/bash
mkdir -p gpurun_out
for cfg in "20,20,20 256 256" "10,10,10 512 512"; do
  set -- $cfg
  echo "=== sk trace DHW=$1 $2->$3"
  DHW=$1 CIN=$2 COUT=$3 timeout 300 python scripts/conv_trace_sk.py 2>&1 | tail -22
done
