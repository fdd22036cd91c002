#!/usr/bin/env bash
# Round 2: the bench on N GPUs of one box exactly as the driver launches it (torchrun, one rank per GPU), + the reference arm.
N=${1:-2}
TAG=${2:-r2m}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 \
    > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
cat gpurun_out/${TAG}_bench_${N}gpu.json; tail -5 gpurun_out/${TAG}_bench_${N}gpu.err
timeout 600 python -m pytest tests/test_gpu_plugin.py -m gpu -q -s -p no:cacheprovider -k "benchmark_app_full_size or sharded" > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "passed|failed|FPS" gpurun_out/${TAG}_pytest.log | tail
