#!/bin/bash
# concurrency stress: 1 / 2 / 4 parallel sub-batch streams of the 32g and 16g denoising step, many graph replays
# (this is what exposed the p_full phase-aliasing deadlock of the one-query-tile attention kernel)
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 600 python scripts/ab_streams.py 6 > gpurun_out/stress_streams_$i.log 2>&1; rc=$?
  echo "run $i exit $rc: $(grep -c 'ms/step' gpurun_out/stress_streams_$i.log) timings, $(grep -c 'timeout\|error\|Error' gpurun_out/stress_streams_$i.log) errors"
done
grep "ms/step" gpurun_out/stress_streams_1.log | tail -18
