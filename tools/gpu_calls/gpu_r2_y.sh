#!/usr/bin/env bash
# experiment: upper bound of cross-kernel overlap - the tcgen05 kernels skip the programmatic-launch wait (racy, results wrong by construction)
TAG=${1:-r2y}
mkdir -p gpurun_out
export B200_TUNE_FILE=$PWD/gpurun_out/${TAG}_tune.txt; rm -f $B200_TUNE_FILE
timeout 900 python tools/gpu_r50_ab.py "" "B200_DEBUG_TC3=4" > gpurun_out/${TAG}_ab.log 2>&1
grep -E "^\[|=====" gpurun_out/${TAG}_ab.log
AB_BATCH=32 timeout 900 python tools/gpu_r50_ab.py "" "B200_DEBUG_TC3=4" > gpurun_out/${TAG}_ab32.log 2>&1
grep -E "^\[|=====" gpurun_out/${TAG}_ab32.log
