#!/bin/bash
# grouped (flat) attention MatMuls: parity test + BERT-base full-size timing with the per-kernel table, against the batched launch
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_membound.py tests/test_gpu_golden.py -x -q -m gpu 2>&1 | tail -5
timeout 400 python tools/gpu_configs.py bert > gpurun_out/q_bert_flat.log 2>&1
B200_MATMUL_FLAT=0 timeout 400 python tools/gpu_configs.py bert > gpurun_out/q_bert_batched.log 2>&1
grep -E "ms|rel|MatMul" gpurun_out/q_bert_flat.log | head -30
echo ---- batched
grep -E "ms/|rel" gpurun_out/q_bert_batched.log | head
