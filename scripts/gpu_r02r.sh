#!/bin/bash
# round 2, call R: full GPU suite + smoke + default bench (cfg1 / configs[3] / configs[4] legs) + reference arm
mkdir -p gpurun_out
echo "=== full gpu suite"
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
echo "=== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "=== bench"
/usr/bin/time -v timeout 1200 python bench.py > gpurun_out/r_bench.log 2> gpurun_out/r_bench.err; echo "exit $?"
grep -E "Elapsed|Maximum resident" gpurun_out/r_bench.err
python scripts/show_bench.py gpurun_out/r_bench.log 2>&1 | head -40
