#!/usr/bin/env bash
# GPU call E: ncu full capture (with source) of pair-kernel launches 15..30 of one ResNet-50 b256 pass
TAG=${1:-e}
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_tc3 -s 15 -c 16 -o gpurun_out/${TAG}_prof \
    python tools/gpu_r50_ab.py --one ncu > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
rm -f gpurun_out/ab_ncu.npy
