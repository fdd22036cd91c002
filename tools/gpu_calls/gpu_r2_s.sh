#!/usr/bin/env bash
# ncu --set full (with SASS stall sampling) of the ResNet-50 b256 layers furthest below their roofline: stem, stage-2 block, first
# stage-3 block, stage 5.  Also re-checks the attention kernel after the row-sum change (tests + BERT config).
TAG=${1:-r2s}
mkdir -p gpurun_out
export B200_TUNE_FILE=$PWD/gpurun_out/${TAG}_tune.txt; rm -f $B200_TUNE_FILE
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-latency --no-benchmark-app > /dev/null 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_tc --launch-skip 0 --launch-count 15 -f -o gpurun_out/${TAG}_a \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-latency --no-benchmark-app > gpurun_out/${TAG}_ncu_a.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_tc --launch-skip 40 --launch-count 9 -f -o gpurun_out/${TAG}_b \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-latency --no-benchmark-app > gpurun_out/${TAG}_ncu_b.log 2>&1
ls -la gpurun_out/${TAG}_*
timeout 600 python -m pytest tests/test_gpu_membound.py tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider -k "attention or bert" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/gpu_configs.py bert > gpurun_out/${TAG}_configs.log 2>&1; grep -E "batch|us " gpurun_out/${TAG}_configs.log | head -14
