#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" >> gpurun_out/$name.log; tail -n 60 gpurun_out/$name.log | cut -c1-400; return $rc; }
TOPN=45 run prof16 python scripts/profile_models.py 16g 36,44,28 8
TOPN=45 run prof32 python scripts/profile_models.py 32g 20 8
run big python scripts/big_configs.py all
run sample python bench.py --workload sample_32g --T 200 --steps 1 --warmup 1
