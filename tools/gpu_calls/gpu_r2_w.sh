#!/usr/bin/env bash
# constant operands fetched before the programmatic-launch wait, .read wait at exit, column-wise global average pool: tests + R50 A/B + configs
TAG=${1:-r2w}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tc3.py tests/test_gpu_ops.py tests/test_gpu_golden.py tests/test_gpu_membound.py tests/test_gpu_configs.py -m gpu -q -x -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
export B200_TUNE_FILE=$PWD/gpurun_out/${TAG}_tune.txt; rm -f $B200_TUNE_FILE
timeout 1500 python tools/gpu_r50_ab.py "" "B200_PDL=0" > gpurun_out/${TAG}_ab.log 2>&1
grep -E "^\[|=====" gpurun_out/${TAG}_ab.log
timeout 600 python tools/gpu_configs.py mobilenet ssd bert > gpurun_out/${TAG}_configs.log 2>&1; grep -E "batch" gpurun_out/${TAG}_configs.log | head -12
timeout 300 python tools/gpu_b1.py > gpurun_out/${TAG}_b1.log 2>&1; tail -8 gpurun_out/${TAG}_b1.log
