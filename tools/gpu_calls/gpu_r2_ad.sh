#!/usr/bin/env bash
# fused inverted-residual block kernel (mbconv.cu): unit tests (each under a timeout: a barrier bug would hang), MobileNet config tests, timing
TAG=${1:-r2ad}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mbconv.py -m gpu -q -p no:cacheprovider --timeout 60 > gpurun_out/${TAG}_pytest_mb.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_mb.log
grep -E "passed|failed|FAILED|Error|rel_l2|Timeout|exit" gpurun_out/${TAG}_pytest_mb.log | head -40
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -p no:cacheprovider -k mobilenet --timeout 200 > gpurun_out/${TAG}_pytest_cfg.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_cfg.log
tail -3 gpurun_out/${TAG}_pytest_cfg.log
B200_DEBUG_MBCONV=1 timeout 300 python tools/gpu_configs.py mobilenet > gpurun_out/${TAG}_configs.log 2>&1; grep -E "batch|us " gpurun_out/${TAG}_configs.log | head -40
grep "b200 mbconv" gpurun_out/${TAG}_configs.log | sort -u | head -20
