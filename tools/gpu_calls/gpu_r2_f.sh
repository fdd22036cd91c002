#!/usr/bin/env bash
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_membound.py tests/test_gpu_ops.py tests/test_gpu_configs.py tests/test_gpu_golden.py -m gpu -q -p no:cacheprovider \
    -k "attention or detection or detout or depthwise or first_convolution or bert or ssd or mobilenet or net_tiny" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "passed|failed|FAILED|ERROR|Error" gpurun_out/${TAG}_pytest.log | tail -20
timeout 600 python tools/gpu_configs.py mobilenet bert ssd > gpurun_out/${TAG}_configs.log 2>&1; grep -E "batch|us " gpurun_out/${TAG}_configs.log | head -60
