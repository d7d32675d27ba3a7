#!/bin/bash
# round 2, call C: polynomial-exp2 offload sweep of the head-dim-64 attention kernel, fused slab GroupNorm v2
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err; rc=$?; echo "exit $rc" >> gpurun_out/$name.err; tail -n 8 gpurun_out/$name.log | cut -c1-300; tail -n 2 gpurun_out/$name.err | cut -c1-300; return $rc; }
for m in 1 2 3 4 5 1 4; do
  TMO=300 B200LDM_ATTN_PACKED=$m run attn_mode${m} python scripts/attn_bench.py
done
for m in 3 5; do
  TMO=600 B200LDM_ATTN_PACKED=$m run t_attn_mode${m} python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "attention"
done
TMO=600 run tc_slabs python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -x -q -m gpu -p no:cacheprovider -k "split_k or sampling or unet_32g or cfg1"
export B200LDM_CONV_TUNE_FILE=$PWD/gpurun_out/conv_tune_r02b.txt
TMO=300 run step32c_plain python scripts/ncu_step.py 32g 20 8
TMO=600 run step32c_ncu ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02c_launches_32g_step.csv python scripts/ncu_step.py 32g 20 8
unset B200LDM_CONV_TUNE_FILE
TMO=600 run bench_c_defer python bench.py --workload sample_32g --T 200 --no-cpu-baseline
