#!/usr/bin/env bash
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_membound.py tests/test_gpu_configs.py -m gpu -q -s -p no:cacheprovider -k "attention or bert" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "passed|failed|FAILED|ERROR|rel-L2|bert-base|Error|timeout" gpurun_out/${TAG}_pytest.log | tail -20
timeout 300 python tools/gpu_configs.py bert > gpurun_out/${TAG}_configs.log 2>&1; grep -E "batch|us " gpurun_out/${TAG}_configs.log | head -16
