#!/bin/bash
# First-contact run on the B200 box: every kernel family in its own process so one CUDA fault does not mask the rest.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
PT="python -m pytest -q --timeout 300 -m gpu -p no:cacheprovider"
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n 25 gpurun_out/$name.log; }
run k_simt   $PT tests/test_gpu_kernels.py -k "not tcgen05 and not tc_bf16"
run k_tc_one $PT tests/test_gpu_kernels.py -k "tcgen05 and k3_basic"
run k_tc_all $PT tests/test_gpu_kernels.py -k "tcgen05 or tc_bf16"
run m_fp32   $PT tests/test_gpu_models.py -k "fp32 or call" -s
run m_bf16   $PT tests/test_gpu_models.py -k "bf16 or graph" -s
run smoke    python __graft_entry__.py smoke
