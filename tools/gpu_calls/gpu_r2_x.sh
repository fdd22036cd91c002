#!/usr/bin/env bash
# experiment: global batch 256 as 1 / 2 / 4 independent sub-batches on concurrent streams of one GPU
TAG=${1:-r2x}
mkdir -p gpurun_out
timeout 900 python tools/gpu_two_stream.py 1 2 4 > gpurun_out/${TAG}_streams.log 2>&1; grep -E "^\[" gpurun_out/${TAG}_streams.log
B200_PDL=0 timeout 900 python tools/gpu_two_stream.py 2 4 > gpurun_out/${TAG}_streams_nopdl.log 2>&1; grep -E "^\[" gpurun_out/${TAG}_streams_nopdl.log
