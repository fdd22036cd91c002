#!/usr/bin/env bash
# ncu full capture (with source) of chosen pair-kernel launches of one all-pair-kernel ResNet-50 b256 pass: usage gpu_call_p.sh <tag> <skip> <count>
TAG=${1:-p}; SKIP=${2:-0}; CNT=${3:-1}
mkdir -p gpurun_out
B200_AUTOTUNE=0 B200_TC3=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_tc3 -s $SKIP -c $CNT -o gpurun_out/${TAG}_prof \
    python tools/gpu_r50_ab.py --one ncu > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
rm -f gpurun_out/ab_ncu.npy
