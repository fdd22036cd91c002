#!/usr/bin/env bash
# Round 2, GPU call C: new kernels' tests, MobileNet / BERT / SSD full-size timing, ncu --set full on the MobileNet CUDA-core kernels.
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_membound.py tests/test_gpu_ops.py tests/test_gpu_configs.py tests/test_gpu_plugin.py -m gpu -q -s -p no:cacheprovider \
    -k "attention or detection or depthwise or first_convolution or bert or ssd or user_memory or export or sharded" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "passed|failed|FAILED|ERROR|rel-L2|ssd300|bert-base" gpurun_out/${TAG}_pytest.log | tail -40
timeout 600 python tools/gpu_configs.py mobilenet bert ssd > gpurun_out/${TAG}_configs.log 2>&1; grep -E "batch|us " gpurun_out/${TAG}_configs.log | head -60
cat > /tmp/mb_run.py <<'PY'
import sys
import numpy as np
sys.path.insert(0, ".")
from dldt_b200 import Core, zoo
net = Core().load_network(zoo.mobilenet_v2(batch=64), "B200", {"B200_INFERENCE_PRECISION": "BF16"})
req = net.create_infer_request()
req.set_blob("data", np.random.default_rng(0).uniform(0, 1, (64, 3, 224, 224)).astype(np.float32))
req.infer(); req.infer()
PY
B200_AUTOTUNE=0 B200_NO_CUDA_GRAPH=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"dwconv3x3_strip|conv_small_cin" -s 18 -c 8 \
    -o gpurun_out/${TAG}_mb python /tmp/mb_run.py > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}_mb.ncu-rep --page raw --csv > gpurun_out/${TAG}_mb_raw.csv 2>/dev/null
python tools/ncu_keymetrics.py gpurun_out/${TAG}_mb_raw.csv
