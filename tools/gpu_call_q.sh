#!/bin/bash
# exact-width packed MVN: parity tests + BERT-base full-size timing against the guarded kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_membound.py tests/test_gpu_golden.py tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -6
timeout 400 python tools/gpu_configs.py bert > gpurun_out/q_bert_mvn.log 2>&1
grep -E "ms/|rel" gpurun_out/q_bert_mvn.log | head; grep -E "^ +[0-9.]+ us +[0-9.]+% x" gpurun_out/q_bert_mvn.log | grep -E "MVN|Softmax"
B200_MVN_EXACT=0 timeout 400 python tools/gpu_configs.py bert > gpurun_out/q_bert_mvn0.log 2>&1
grep -E "ms/" gpurun_out/q_bert_mvn0.log | head; grep -E "^ +[0-9.]+ us +[0-9.]+% x" gpurun_out/q_bert_mvn0.log | grep -E "MVN|Softmax"
