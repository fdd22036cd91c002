#!/usr/bin/env bash
# shift/mask bf16 unpack (all tcgen05 epilogues + mbconv), fp32 depthwise weights in shared memory: full GPU suite, MobileNet and R50 timing
TAG=${1:-r2an}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -m gpu -q -x -p no:cacheprovider --timeout 300 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/gpu_configs.py mobilenet > gpurun_out/${TAG}_configs.log 2>&1; grep -E "batch|us " gpurun_out/${TAG}_configs.log | head -18
timeout 600 python tools/gpu_r50_ab.py "" > gpurun_out/${TAG}_ab.log 2>&1; grep -E "^\[" gpurun_out/${TAG}_ab.log
