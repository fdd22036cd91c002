#!/usr/bin/env bash
# GPU call: tc3 probe (per-case subprocesses) + R50 b256 per-layer table with tc2 only and with tc3 auto
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
timeout 1500 python tools/gpu_tc3_probe.py 2>&1 | tee gpurun_out/c1_tc3_probe.log
PROBE_SKIP_CONV=1 PROBE_SKIP_FP32=1 B200_TC3=0 timeout 300 python tools/gpu_probe.py > gpurun_out/c1_r50_tc2.log 2>&1; grep "img/s" gpurun_out/c1_r50_tc2.log
PROBE_SKIP_CONV=1 PROBE_SKIP_FP32=1 timeout 300 python tools/gpu_probe.py > gpurun_out/c1_r50_tc3.log 2>&1; grep "img/s\|Error\|error" gpurun_out/c1_r50_tc3.log | head
PROBE_SKIP_CONV=1 PROBE_SKIP_FP32=1 B200_TC3_ROWS=0 timeout 300 python tools/gpu_probe.py > gpurun_out/c1_r50_tc3_norows.log 2>&1; grep "img/s\|Error\|error" gpurun_out/c1_r50_tc3_norows.log | head
