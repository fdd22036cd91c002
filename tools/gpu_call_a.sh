#!/usr/bin/env bash
# GPU call A: CTA-pair kernel unit probe, parity suite with the pair kernel off (known-good baseline), ResNet-50 A/B of the knobs.
TAG=${1:-a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 700 python tools/gpu_tc3_probe.py > gpurun_out/${TAG}_tc3_probe.log 2>&1; cat gpurun_out/${TAG}_tc3_probe.log
B200_TC3=0 timeout 1000 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/${TAG}_pytest_tc3off.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_tc3off.log
tail -12 gpurun_out/${TAG}_pytest_tc3off.log
timeout 1200 python tools/gpu_r50_ab.py "B200_TC3=0" "" "B200_TC3_HALO=1" "B200_TC3_BASEOFF=1" "B200_TC3_HALO=1 B200_TC3_BASEOFF=1" > gpurun_out/${TAG}_ab.log 2>&1
grep -E "^\[|=====|exit|TIMEOUT" gpurun_out/${TAG}_ab.log
