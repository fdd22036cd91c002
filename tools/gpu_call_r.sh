#!/bin/bash
# ncu --set full on the attention softmax alone (scores [16,12,384,384] bf16), one row per warp vs four rows per warp
mkdir -p gpurun_out
cat > /tmp/sm_run.py <<'PY'
import sys
import numpy as np
sys.path.insert(0, ".")
from dldt_b200 import Core, Function
f = Function("attn")
x = f.parameter((16, 12, 384, 384), name="scores")
m = f.parameter((16, 1, 1, 384), name="mask")
f.result(f.softmax(f.add(f.multiply(x, f.constant(np.float32(0.125))), m), 3))
net = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": "BF16"})
req = net.create_infer_request()
rng = np.random.default_rng(0)
req.set_blob("scores", rng.standard_normal((16, 12, 384, 384)).astype(np.float32))
req.set_blob("mask", np.zeros((16, 1, 1, 384), np.float32))
req.infer(); req.infer()
print("us/launch", req.time_launches(20) / 20 * 1e3)
PY
for rpw in 1 4; do
  B200_SOFTMAX_RPW=$rpw python /tmp/sm_run.py 2>&1 | tail -1
  B200_SOFTMAX_RPW=$rpw B200_NO_CUDA_GRAPH=1 timeout 300 ncu --set full --clock-control none -k regex:"softmax_warp" -s 1 -c 1 -o gpurun_out/r_softmax_rpw$rpw -f python /tmp/sm_run.py > gpurun_out/r_softmax_rpw$rpw.log 2>&1
  tail -1 gpurun_out/r_softmax_rpw$rpw.log
done
