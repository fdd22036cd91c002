#!/usr/bin/env bash
# ncu --set full of the fused MobileNet block kernel (ir1_1: 56x56x24 -> 144 -> 24, residual) + per-source-line stall table
TAG=${1:-r2ai}
mkdir -p gpurun_out /tmp/cub
cat > /tmp/mb_run.py <<'PY'
import sys
import numpy as np
sys.path.insert(0, ".")
from dldt_b200 import Core, zoo
net = Core().load_network(zoo.mobilenet_v2(batch=64), "B200", {"B200_INFERENCE_PRECISION": "BF16"})
req = net.create_infer_request()
req.set_blob("data", np.random.default_rng(0).uniform(0, 1, (64, 3, 224, 224)).astype(np.float32))
req.infer()
PY
B200_NO_CUDA_GRAPH=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:mbconv --launch-skip 0 --launch-count 3 -f -o gpurun_out/${TAG}_mb python /tmp/mb_run.py > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}_mb.ncu-rep --page raw --csv > gpurun_out/${TAG}_mb_raw.csv 2>/dev/null
python tools/ncu_keymetrics.py gpurun_out/${TAG}_mb_raw.csv | cut -c1-500
(cd /tmp/cub && cuobjdump -xelf all $OLDPWD/dldt_b200/lib/libb200kernels.so > /dev/null 2>&1)
python tools/stall_lines.py gpurun_out/${TAG}_mb.ncu-rep 0 /tmp/cub/mbconv.sm_100a.cubin 13__nv_bfloat16Li1ELb0E 8 > gpurun_out/${TAG}_stalls_ir0_0.txt 2>&1
python tools/stall_lines.py gpurun_out/${TAG}_mb.ncu-rep 2 /tmp/cub/mbconv.sm_100a.cubin 13__nv_bfloat16Li1ELb1E 8 > gpurun_out/${TAG}_stalls_ir1_1.txt 2>&1
head -40 gpurun_out/${TAG}_stalls_ir0_0.txt | cut -c1-200
ls -la gpurun_out
