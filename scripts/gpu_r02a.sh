#!/bin/bash
# round 2, call A: full GPU test suite, smoke, the new bench line, launch lists of the 32g / 16g steps, and the
# whole-batch 1000-step bf16-vs-fp32-engine trajectory
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err; rc=$?; echo "exit $rc" >> gpurun_out/$name.err; tail -n 4 gpurun_out/$name.log | cut -c1-700; tail -n 3 gpurun_out/$name.err | cut -c1-400; return $rc; }
TMO=1500 run t_gpu python -m pytest tests -x -q -m gpu -p no:cacheprovider --durations=15
TMO=300 run smoke python __graft_entry__.py smoke
TMO=900 run bench python bench.py
TMO=600 run bench_ref python bench.py --impl reference --steps 2 --warmup 1
export B200LDM_CONV_TUNE_FILE=$PWD/gpurun_out/conv_tune_r02a.txt
rm -f $B200LDM_CONV_TUNE_FILE
TMO=300 run step32_plain python scripts/ncu_step.py 32g 20 8
TMO=600 run step32_ncu ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02a_launches_32g_step.csv python scripts/ncu_step.py 32g 20 8
TMO=300 run step16_plain python scripts/ncu_step.py 16g 36,44,28 8
TMO=600 run step16_ncu ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02a_launches_16g_step.csv python scripts/ncu_step.py 16g 36,44,28 8
TMO=900 run lp_cfg3_b8_fp32_1000 python scripts/long_parity.py cfg3 --latent 20 --steps 1000 --batch-fp32 8 --every 100
