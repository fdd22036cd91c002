#!/usr/bin/env bash
TAG=${1:-r2h}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_configs.py tests/test_gpu_golden.py -m gpu -q -p no:cacheprovider -k "detection or detout or ssd" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "passed|failed|FAILED|ERROR|Error" gpurun_out/${TAG}_pytest.log | tail -10
timeout 300 python tools/gpu_layers.py mobilenet > gpurun_out/${TAG}_mobilenet_layers.log 2>&1; cat gpurun_out/${TAG}_mobilenet_layers.log
timeout 300 python tools/gpu_layers.py ssd 2>&1 | grep -E "ms/step|Detection|sum of" 
