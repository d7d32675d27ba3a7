#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/stress_profile.log
for i in $(seq 1 ${1:-6}); do
  python scripts/profile_models.py 32g 20 8 > gpurun_out/sp_$i.log 2>&1; rc=$?
  echo "run $i rc=$rc $(grep -c 'decode eager' gpurun_out/sp_$i.log)" >> gpurun_out/stress_profile.log
  if [ $rc -ne 0 ]; then grep -n "timeout\|Error\|error" gpurun_out/sp_$i.log | head -5 >> gpurun_out/stress_profile.log; fi
done
cat gpurun_out/stress_profile.log
