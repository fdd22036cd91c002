#!/usr/bin/env bash
# stem + max-pool fusion: tests, R50 A/B (fusion on / off), full-network parity tests
TAG=${1:-r2aa}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc3.py -m gpu -q -x -p no:cacheprovider -k "stem or s2d" > gpurun_out/${TAG}_pytest_stem.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_stem.log
tail -15 gpurun_out/${TAG}_pytest_stem.log
export B200_TUNE_FILE=$PWD/gpurun_out/${TAG}_tune.txt; rm -f $B200_TUNE_FILE
timeout 900 python tools/gpu_r50_ab.py "B200_FUSE_POOL=0" "" > gpurun_out/${TAG}_ab.log 2>&1
grep -E "^\[|=====|conv1|pool1|zero" gpurun_out/${TAG}_ab.log
unset B200_TUNE_FILE
timeout 1500 python -m pytest tests/ -m gpu -q -x -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
