#!/bin/bash
# round 2, call W: full GPU suite + default bench with stream-K among the candidates
mkdir -p gpurun_out
echo "=== full gpu suite"
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
echo "=== bench"
t0=$(date +%s)
timeout 1200 python bench.py > gpurun_out/w_bench.log 2> gpurun_out/w_bench.err; echo "exit $? after $(( $(date +%s) - t0 )) s"
tail -2 gpurun_out/w_bench.err
python scripts/show_bench.py gpurun_out/w_bench.log 2>&1 | head -40
