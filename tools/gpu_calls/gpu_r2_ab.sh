#!/usr/bin/env bash
# s2d stem for 3x3/5x5 first convolutions (MobileNet), min-A-stages tweak: tests + configs + R50 check
TAG=${1:-r2ac}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tc3.py tests/test_gpu_membound.py tests/test_gpu_configs.py tests/test_gpu_golden.py -m gpu -q -x -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -8 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/gpu_configs.py mobilenet > gpurun_out/${TAG}_configs.log 2>&1; grep -E "batch|us " gpurun_out/${TAG}_configs.log | head -16
timeout 900 python tools/gpu_r50_ab.py "" > gpurun_out/${TAG}_ab.log 2>&1
grep -E "^\[|=====|res3_0_b" gpurun_out/${TAG}_ab.log
