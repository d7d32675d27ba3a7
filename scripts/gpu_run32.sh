#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --timeout 600 -m gpu -p no:cacheprovider"
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" >> gpurun_out/$name.log; tail -n 8 gpurun_out/$name.log | cut -c1-300; return $rc; }
run t_all $PT tests
run smoke python __graft_entry__.py smoke
run big python scripts/big_configs.py all
