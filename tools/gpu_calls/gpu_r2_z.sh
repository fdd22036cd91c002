#!/usr/bin/env bash
# chained 1x1 kernel (conv_tc3f) against the current pair kernel on ResNet-50 b256
TAG=${1:-r2z}
mkdir -p gpurun_out
timeout 900 python tools/gpu_r50_ab.py "" "B200_FUSE_NEXT=1" > gpurun_out/${TAG}_ab.log 2>&1
grep -E "^\[|=====" gpurun_out/${TAG}_ab.log
