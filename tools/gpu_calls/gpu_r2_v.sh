#!/usr/bin/env bash
# tests of the single-CTA kernel's straight-line epilogue, then timing experiments on ResNet-50 b256 (debug knobs: results wrong by construction)
TAG=${1:-r2v}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tc3.py tests/test_gpu_ops.py tests/test_gpu_golden.py tests/test_gpu_membound.py -m gpu -q -x -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
export B200_TUNE_FILE=$PWD/gpurun_out/${TAG}_tune.txt; rm -f $B200_TUNE_FILE
timeout 1500 python tools/gpu_r50_ab.py "" "B200_DEBUG_TC3=1" "B200_DEBUG_TC3=3" "B200_DEBUG_NO_RES=1" "B200_DEBUG_NO_RES=1 B200_DEBUG_TC3=3" "B200_TC3_EVICT=0" "B200_TC3_WSN=0" > gpurun_out/${TAG}_ab.log 2>&1
grep -E "^\[|=====" gpurun_out/${TAG}_ab.log
timeout 600 python tools/gpu_configs.py mobilenet > gpurun_out/${TAG}_configs.log 2>&1; grep -E "batch|us " gpurun_out/${TAG}_configs.log | head -12
