#!/usr/bin/env bash
# pure kernel durations (ncu) of one MobileNetV2 b64 step vs the graph-replay step time: how much of the step is launch gaps?
TAG=${1:-r2i}
mkdir -p gpurun_out
cat > /tmp/mb_run.py <<'PY'
import sys
import numpy as np
sys.path.insert(0, ".")
from dldt_b200 import Core, zoo
net = Core().load_network(zoo.mobilenet_v2(batch=64), "B200", {"B200_INFERENCE_PRECISION": "BF16"})
req = net.create_infer_request()
req.set_blob("data", np.random.default_rng(0).uniform(0, 1, (64, 3, 224, 224)).astype(np.float32))
req.infer(); req.infer()
print("graph ms", min(req.time_launches(20) / 20 for _ in range(3)))
PY
python /tmp/mb_run.py 2>&1 | tail -1
B200_PDL=0 python /tmp/mb_run.py 2>&1 | tail -1
B200_NO_CUDA_GRAPH=1 python /tmp/mb_run.py 2>&1 | tail -1
B200_TUNE_FILE=/tmp/tune.txt python /tmp/mb_run.py > /dev/null 2>&1
B200_TUNE_FILE=/tmp/tune.txt B200_NO_CUDA_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_mb_launches.csv python /tmp/mb_run.py > gpurun_out/${TAG}_ncu.log 2>&1
python tools/summarise_ncu.py launches gpurun_out/${TAG}_mb_launches.csv
python - <<'PY'
import csv
rows=[l for l in open("gpurun_out/r2i_mb_launches.csv") if l.startswith('"')]
r=list(csv.DictReader(rows))
d=[float(x["Metric Value"].replace(",",""))/1e3 for x in r if x["Metric Name"]=="gpu__time_duration.sum"]
n=len(d)//3 if len(d)>=168 else len(d)
print("launches", len(d), "last step kernels:", [round(v,1) for v in d[-56:]])
print("last step sum us", sum(d[-56:]))
PY
