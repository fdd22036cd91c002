#!/usr/bin/env bash
# straight-line slab epilogue + big resident weights in halo mode: GPU tests, then A/B on ResNet-50 b256, MobileNet / SSD / BERT configs
TAG=${1:-r2u}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tc3.py tests/test_gpu_ops.py tests/test_gpu_golden.py tests/test_gpu_membound.py -m gpu -q -x -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 900 python tools/gpu_r50_ab.py "B200_TC3_WS_BIG=0" "" > gpurun_out/${TAG}_ab.log 2>&1
grep -E "^\[|=====" gpurun_out/${TAG}_ab.log
timeout 600 python tools/gpu_configs.py mobilenet ssd bert > gpurun_out/${TAG}_configs.log 2>&1; grep -E "batch" gpurun_out/${TAG}_configs.log
