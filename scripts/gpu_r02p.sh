#!/bin/bash
# round 2, call P: finalize-with-statistics for split-K convs, forced-split test, 32g step again, default bench with the
# configs[3] / configs[4] legs
mkdir -p gpurun_out
echo "=== conv kernel tests"
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "conv or split or gn or norm" 2>&1 | tail -4
echo "=== 32g step"
B200LDM_CONV_AUTOTUNE_LOG=1 PROFILE_STREAMS=2 TOPN=40 timeout 600 python scripts/profile_models.py 32g 20 8 > gpurun_out/p_32g.log 2> gpurun_out/p_32g.err
grep -E "graph replay|profiled kernels|decode eager|gn_|conv3d  " gpurun_out/p_32g.log | head -20
echo "=== model tests"
timeout 900 python -m pytest tests/test_gpu_models.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
echo "=== bench"
timeout 1200 python bench.py > gpurun_out/p_bench.log 2> gpurun_out/p_bench.err; echo "exit $?"
tail -3 gpurun_out/p_bench.err
python scripts/show_bench.py gpurun_out/p_bench.log 2>/dev/null | head -40
