#!/bin/bash
# Long-trajectory / full-size parity of the BASELINE.json configurations against the CPU oracle (one gpurun call):
#   gpurun --timeout 2700 -- 'bash scripts/gpu_long_parity.sh'
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err; rc=$?; echo "exit $rc" >> gpurun_out/$name.err; tail -n 3 gpurun_out/$name.log | cut -c1-1500; tail -n 3 gpurun_out/$name.err | cut -c1-600; return $rc; }
nproc; free -g | head -2
TMO=300  run lp_quick  python scripts/long_parity.py cfg3 --latent 16 --steps 20 --every 10
TMO=300  run lp_cfg5_96  python scripts/long_parity.py cfg5 --edge 96
TMO=1500 run lp_cfg3_20 python scripts/long_parity.py cfg3 --latent 20 --steps 1000 --fixture gpurun_out/cfg3_ddpm1000_20.pt
TMO=300  run lp_cfg5_128 python scripts/long_parity.py cfg5 --edge 128
TMO=900  run lp_cfg4_40 python scripts/long_parity.py cfg4 --latent 40 --steps 50
TMO=600  run lp_cfg3_b8_fp32 python scripts/long_parity.py cfg3 --latent 20 --steps 100 --batch-fp32 8
