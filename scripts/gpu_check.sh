#!/bin/bash
# What the driver runs at round end, in one gpurun call:  gpurun --timeout 1800 -- 'bash scripts/gpu_check.sh'
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" >> gpurun_out/$name.log; tail -n 6 gpurun_out/$name.log | cut -c1-400; return $rc; }
run t_gpu python -m pytest tests -x -q -m gpu -p no:cacheprovider
run smoke python __graft_entry__.py smoke
run bench python bench.py
run bench_ref python bench.py --impl reference --steps 1 --warmup 1
