#!/bin/bash
# ncu --set full (SASS stall sampling) of the FFN-up FullyConnected with the GELU epilogue
mkdir -p gpurun_out
cat > /tmp/gelu_run.py <<'PY'
import sys
import numpy as np
sys.path.insert(0, ".")
from dldt_b200 import Core, Function
rng = np.random.default_rng(0)
f = Function("fc")
x = f.parameter((64, 384, 768), name="x")
w = (rng.standard_normal((3072, 768)) / 27.7).astype(np.float32)
y = f.gelu(f.add(f.matmul(x, w, transpose_b=True), f.constant(rng.uniform(-1, 1, (3072,)).astype(np.float32))), "ERF")
f.result(y)
net = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": "BF16"})
req = net.create_infer_request()
req.infer(); req.infer()
PY
B200_AUTOTUNE=0 B200_NO_CUDA_GRAPH=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_tc3 -s 1 -c 1 -o gpurun_out/s_gelu -f python /tmp/gelu_run.py > gpurun_out/s_gelu.log 2>&1
tail -2 gpurun_out/s_gelu.log
