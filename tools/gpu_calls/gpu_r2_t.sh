#!/usr/bin/env bash
# (1) pair-kernel tests + A/B of the TMEM accumulator ring / early accumulator hand-back on ResNet-50 b256;
# (2) ncu --set full (SASS stall sampling) of selected ResNet-50 b256 launches, reports kept small (<= 64 MiB per call)
TAG=${1:-r2t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc3.py -m gpu -q -x -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 900 python tools/gpu_r50_ab.py "B200_TC3_NACC=2 B200_TC3_EARLY=0" "B200_TC3_NACC=2" "B200_TC3_EARLY=0" "" > gpurun_out/${TAG}_ab.log 2>&1
grep -E "^\[|=====" gpurun_out/${TAG}_ab.log
export B200_TUNE_FILE=$PWD/gpurun_out/${TAG}_tune.txt; rm -f $B200_TUNE_FILE
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-latency --no-benchmark-app > /dev/null 2>&1
for R in "0 3" "5 2" "14 2" "26 2" "44 2"; do
  set -- $R
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tc --launch-skip $1 --launch-count $2 -f -o gpurun_out/${TAG}_l$1 \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-latency --no-benchmark-app > gpurun_out/${TAG}_ncu_l$1.log 2>&1
  ncu -i gpurun_out/${TAG}_l$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_l$1_raw.csv 2>/dev/null
  for k in $(seq 0 $(($2 - 1))); do python tools/sass_stalls.py gpurun_out/${TAG}_l$1.ncu-rep $k 40 > gpurun_out/${TAG}_l$1_k${k}_stalls.txt 2>&1; done
done
du -sh gpurun_out; ls -la gpurun_out | head -40
# keep the call under the 64 MiB pull limit: drop the largest reports first
while [ $(du -sm gpurun_out | cut -f1) -ge 60 ]; do f=$(ls -S gpurun_out/*.ncu-rep | head -1); echo "dropping $f"; rm -f $f; done
