#!/usr/bin/env bash
# GPU call C: pair-kernel probe (correctness) + ResNet-50 A/B of the knobs given as arguments
TAG=${1:-c}; shift
mkdir -p gpurun_out
timeout 700 python tools/gpu_tc3_probe.py > gpurun_out/${TAG}_tc3_probe.log 2>&1; grep -c "OK " gpurun_out/${TAG}_tc3_probe.log; grep -v "OK " gpurun_out/${TAG}_tc3_probe.log
timeout 1200 python tools/gpu_r50_ab.py "$@" > gpurun_out/${TAG}_ab.log 2>&1
grep -E "^\[|=====|exit|TIMEOUT" gpurun_out/${TAG}_ab.log
