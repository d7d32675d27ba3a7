#!/bin/bash
# round 2, call B: packed-fp32x2 attention A/B, split-K finalize fused into GroupNorm (tests + 32g launch list), bench
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err; rc=$?; echo "exit $rc" >> gpurun_out/$name.err; tail -n 12 gpurun_out/$name.log | cut -c1-400; tail -n 3 gpurun_out/$name.err | cut -c1-400; return $rc; }
TMO=900 run tb_gpu python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -x -q -m gpu -p no:cacheprovider
TMO=300 run attn_packed1 python scripts/attn_bench.py
TMO=300 B200LDM_ATTN_PACKED=0 run attn_packed0 python scripts/attn_bench.py
TMO=300 run attn_packed1b python scripts/attn_bench.py
export B200LDM_CONV_TUNE_FILE=$PWD/gpurun_out/conv_tune_r02b.txt
rm -f $B200LDM_CONV_TUNE_FILE
TMO=300 run step32b_plain python scripts/ncu_step.py 32g 20 8
TMO=600 run step32b_ncu ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_launches_32g_step.csv python scripts/ncu_step.py 32g 20 8
unset B200LDM_CONV_TUNE_FILE
TMO=900 run bench_b python bench.py
TMO=600 B200LDM_DEFER_FINALIZE=0 run bench_b_nodefer python bench.py --workload sample_32g --T 200 --no-cpu-baseline
TMO=600 run bench_b_defer python bench.py --workload sample_32g --T 200 --no-cpu-baseline
