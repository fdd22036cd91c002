#!/usr/bin/env bash
# One gpurun call: parity tests, probe, bench, ncu launch list, ncu full capture of the conv kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -25 >> gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/gpu_probe.py > gpurun_out/probe.log 2>&1; tail -40 gpurun_out/probe.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 108 -c 14 -o gpurun_out/prof_conv_tc \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
