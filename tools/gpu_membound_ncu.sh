#!/usr/bin/env bash
# ncu evidence for the memory-bound kernel classes: DRAM bytes and duration per launch (MobileNetV2 b64, BERT-base b16, ResNet-50 b256 stem pool)
TAG=${1:-mb}
mkdir -p gpurun_out
cat > /tmp/mb_run.py <<'PY'
import sys
import numpy as np
sys.path.insert(0, ".")
from dldt_b200 import Core, zoo
core = Core()
for name, f, prec in (("mobilenet", zoo.mobilenet_v2(batch=64), "BF16"), ("bert", zoo.bert_base(batch=16), "BF16"), ("resnet", zoo.resnet50(batch=256), "BF16")):
    net = core.load_network(f, "B200", {"B200_INFERENCE_PRECISION": prec, })
    req = net.create_infer_request()
    rng = np.random.default_rng(0)
    for nm, shape, _ in net.inputs:
        if name == "bert":
            req.set_blob(nm, np.ones(shape, np.int32) if "mask" in nm else (np.zeros(shape, np.int32) if "type" in nm else rng.integers(0, 30000, shape, dtype=np.int32)))
        else:
            req.set_blob(nm, rng.uniform(0, 1, shape).astype(np.float32))
    req.infer()
    req.infer()
PY
B200_AUTOTUNE=0 B200_NO_CUDA_GRAPH=1 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:"maxpool3x3s2|dwconv3x3|softmax_warp|mvn_warp|permute_vec16|ingest_s2d|global_avgpool_warp|gather_rows|binary_" -c 400 --csv \
    --log-file gpurun_out/${TAG}_membound.csv python /tmp/mb_run.py > gpurun_out/${TAG}_membound.log 2>&1
tail -2 gpurun_out/${TAG}_membound.log
