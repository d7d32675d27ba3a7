#!/bin/bash
# fresh processes: each one measures its conv kernel choices again, then decodes and samples
mkdir -p gpurun_out; : > gpurun_out/stress_fresh.log
for i in $(seq 1 ${1:-8}); do
  python scripts/stress_decode.py 3 >> gpurun_out/stress_fresh.log 2>&1; echo "run $i rc=$?" >> gpurun_out/stress_fresh.log
done
grep -c "rc=0" gpurun_out/stress_fresh.log; grep -v MISMATCH gpurun_out/stress_fresh.log | grep -v "first decode\|stress done\|rc=0" | head -20
