#!/usr/bin/env bash
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_membound.py tests/test_gpu_ops.py tests/test_gpu_configs.py tests/test_gpu_golden.py -m gpu -q -p no:cacheprovider \
    -k "attention or softmax or bert or ssd or net_tiny" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "passed|failed|FAILED|ERROR|Error" gpurun_out/${TAG}_pytest.log | tail -20
timeout 600 python tools/gpu_configs.py bert ssd > gpurun_out/${TAG}_configs.log 2>&1; grep -E "batch|us " gpurun_out/${TAG}_configs.log | grep -E "batch|Attention|Detection|Softmax" | head
cat > /tmp/ssd_run.py <<'PY'
import sys
import numpy as np
sys.path.insert(0, ".")
from dldt_b200 import Core, zoo
net = Core().load_network(zoo.ssd300_resnet34(batch=32, head_scale=0.02), "B200", {"B200_INFERENCE_PRECISION": "FP16"})
req = net.create_infer_request()
req.set_blob("data", np.random.default_rng(0).uniform(0, 1, (32, 3, 300, 300)).astype(np.float32))
req.infer(); req.infer()
PY
B200_AUTOTUNE=0 B200_NO_CUDA_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"detout|softmax|convert_flat|copy2d" -c 200 --csv \
    --log-file gpurun_out/${TAG}_ssd_launches.csv python /tmp/ssd_run.py > gpurun_out/${TAG}_ncu.log 2>&1
python tools/summarise_ncu.py launches gpurun_out/${TAG}_ssd_launches.csv
