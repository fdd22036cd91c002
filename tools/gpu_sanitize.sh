#!/usr/bin/env bash
# compute-sanitizer memcheck over a slice of the GPU parity tests (every kernel family once, small shapes): out-of-bounds / misaligned
# global, shared and TMEM-adjacent accesses, leaked allocations.  Usage (on the GPU box): tools/gpu_sanitize.sh [pytest -k expression]
set -uo pipefail
cd "$(dirname "$0")/.."
K="${1:-tiny_resnet or tiny_mobilenet or tiny_bert or tiny_ssd or fc_k_not or conv_bias_relu}"
mkdir -p gpurun_out
B200_AUTOTUNE=0 timeout "${SANITIZE_TIMEOUT:-600}" compute-sanitizer --tool memcheck --error-exitcode 86 --print-limit 30 \
  --log-file gpurun_out/sanitizer_memcheck.log \
  python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "$K" -p no:cacheprovider > gpurun_out/sanitizer_pytest.log 2>&1
echo "sanitizer run exit code $?" | tee -a gpurun_out/sanitizer_pytest.log
tail -n 5 gpurun_out/sanitizer_pytest.log
grep -E "ERROR SUMMARY|Invalid|Misaligned|leaked" gpurun_out/sanitizer_memcheck.log | head -20
