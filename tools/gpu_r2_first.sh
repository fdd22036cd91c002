#!/usr/bin/env bash
# Round 2, first GPU call: every -m gpu test (no -x: all failures in one pass), bench + reference arm, full-size configs.
TAG=${1:-r2a}
mkdir -p gpurun_out
rm -f gpurun_out/lowp_errors.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|rel-L2|benchmark_app|ssd300|bert-base" gpurun_out/${TAG}_pytest_gpu.log | tail -60
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 400 python bench.py --impl reference --steps 12 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; cat gpurun_out/${TAG}_bench_ref.json; tail -3 gpurun_out/${TAG}_bench_ref.err
timeout 600 python tools/gpu_configs.py > gpurun_out/${TAG}_configs.log 2>&1; cat gpurun_out/${TAG}_configs.log | head -70
