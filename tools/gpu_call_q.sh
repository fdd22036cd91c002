#!/bin/bash
# MN-major B operand for the context MatMul (no V^T transpose): parity tests + BERT-base full-size timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_membound.py tests/test_gpu_golden.py -x -q -m gpu 2>&1 | tail -12
timeout 400 python tools/gpu_configs.py bert > gpurun_out/q_bert_bmn.log 2>&1
grep -E "ms/|rel" gpurun_out/q_bert_bmn.log | head; grep -E "^ +[0-9.]+ us +[0-9.]+% x" gpurun_out/q_bert_bmn.log
