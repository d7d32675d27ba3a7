#!/bin/bash
# round 2, call L (1 GPU): BASELINE configs[2] strong scaling at N = 1 - all 64 volumes on one GPU - with 8 / 32 / 64 volumes
# per sampling call: does a larger per-GPU batch lift volumes/s/GPU?
mkdir -p gpurun_out
for b in 8 32 64; do
  echo "=== strong N=1 batch $b"
  timeout 900 python bench.py --workload sample_32g --T 1000 --scaling strong --total-volumes 64 --batch $b --steps 1 --warmup 1 > gpurun_out/strong_n1_b$b.log 2> gpurun_out/strong_n1_b$b.err; echo "exit $?"
  python - <<PY
import json
l=[x for x in open("gpurun_out/strong_n1_b$b.log").read().splitlines() if x.startswith("{")]
if l:
    d=json.loads(l[0]); print("batch $b:", round(d["value"],3), "volumes/s", round(d["ms_per_step"]/1e3,2), "s for 64 volumes; frac_executed", round(d["frac_executed"],3), "e2e", round(d["e2e"]["value"],3))
PY
done
