#!/usr/bin/env bash
# ncu full capture (with source) of chained-kernel launches of one ResNet-50 b256 pass: usage gpu_call_f.sh <tag> <skip> <count>
TAG=${1:-f}; SKIP=${2:-4}; CNT=${3:-2}
mkdir -p gpurun_out
B200_AUTOTUNE=0 timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_tc3f -s $SKIP -c $CNT -o gpurun_out/${TAG}_prof \
    python tools/gpu_r50_ab.py --one ncu > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
rm -f gpurun_out/ab_ncu.npy
