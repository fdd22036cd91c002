#!/usr/bin/env bash
# Round 2: the bench on N GPUs of one box exactly as the driver launches it (torchrun, one rank per GPU), + sharded / MULTI benchmark_app runs
N=${1:-2}
TAG=${2:-r2m}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 \
    > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_${N}gpu.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["benchmark_app"], "\nstrong", d["strong"], "\nmirror", d["e2e_python_mirror"]["value"])
PY
tail -3 gpurun_out/${TAG}_bench_${N}gpu.err
# request-count sweep of the batch-sharded mode (one process, all GPUs)
python - <<'PY' > gpurun_out/${TAG}_nireq_sweep.log 2>&1
import os, sys, tempfile
sys.path.insert(0, ".")
import bench
tmp = tempfile.mkdtemp()
xml = bench.write_ir(256, tmp)
for nireq in (2, 3, 4, 6, 8, 12):
    r = bench.benchmark_app(xml, "B200.ALL", 4, extra=["-nireq", str(nireq)])
    print("B200.ALL nireq %2d: %.0f FPS, latency %.2f ms" % (nireq, r["fps"], r["latency_ms"] or -1), flush=True)
PY
cat gpurun_out/${TAG}_nireq_sweep.log
timeout 300 python -m pytest tests/test_gpu_membound.py tests/test_gpu_golden.py -m gpu -q -p no:cacheprovider -k "first_convolution or net_tiny_mobilenet or depthwise" 2>&1 | tail -2
