#!/usr/bin/env bash
# FakeQuantize / Round (C ABI, plugin, golden cases), adaptive hidden-chunk ring in mbconv: full GPU suite + MobileNet timing
TAG=${1:-r2ap}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/ -m gpu -q -x -p no:cacheprovider --timeout 300 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/gpu_configs.py mobilenet > gpurun_out/${TAG}_configs.log 2>&1; grep -E "batch" gpurun_out/${TAG}_configs.log | head -3
