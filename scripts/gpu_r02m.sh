#!/bin/bash
# round 2, call M (N GPUs, torchrun): the default bench line (weak scaling) and BASELINE configs[2] strong scaling
# (64 volumes shared by the GPUs).   gpurun --gpus N -- 'bash scripts/gpu_r02m.sh N'
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "=== default bench N=$N"
timeout 900 $TR bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "exit $?"
grep '^{' gpurun_out/bench_n$N.log | cut -c1-300
echo "=== strong sample_32g N=$N"
timeout 900 $TR bench.py --gpus $N --workload sample_32g --T 1000 --scaling strong --total-volumes 64 --batch $((64 / N)) --steps 1 --warmup 1 > gpurun_out/strong_n$N.log 2> gpurun_out/strong_n$N.err; echo "exit $?"
grep '^{' gpurun_out/strong_n$N.log | cut -c1-300
