#!/usr/bin/env bash
# ncu --set full of the fused attention kernel at the BERT-base b64 layer shape
TAG=${1:-r2e}
mkdir -p gpurun_out
cat > /tmp/attn_run.py <<'PY'
import sys
import numpy as np
sys.path.insert(0, ".")
from dldt_b200 import Core, Function
n, heads, seq = 64, 12, 384
f = Function("attention")
q = f.parameter((n, heads, seq, 64), name="q"); k = f.parameter((n, heads, seq, 64), name="k"); v = f.parameter((n, heads, seq, 64), name="v")
m = f.parameter((n, 1, 1, seq), name="mask")
s = f.add(f.multiply(f.matmul(q, k, transpose_b=True), np.float32(0.125)), m)
c = f.matmul(f.softmax(s, 3), v)
f.result(f.reshape(f.transpose(c, (0, 2, 1, 3)), (n, seq, heads * 64)))
rng = np.random.default_rng(0)
net = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": "BF16"})
req = net.create_infer_request()
for nm in ("q", "k", "v"):
    req.set_blob(nm, rng.standard_normal((n, heads, seq, 64)).astype(np.float32))
req.set_blob("mask", np.zeros((n, 1, 1, seq), np.float32))
req.infer(); req.infer()
print("ms", req.time_launches(20) / 20)
PY
python /tmp/attn_run.py 2>&1 | tail -1
B200_NO_CUDA_GRAPH=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"attention_kernel" -s 1 -c 1 -o gpurun_out/${TAG}_attn python /tmp/attn_run.py > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}_attn.ncu-rep --page raw --csv > gpurun_out/${TAG}_attn_raw.csv 2>/dev/null
python tools/ncu_keymetrics.py gpurun_out/${TAG}_attn_raw.csv
python tools/sass_stalls.py gpurun_out/${TAG}_attn.ncu-rep 0 40 2>&1 | head -70
