#!/bin/bash
# one-switch-per-slab generic epilogue: parity tests, FC epilogue microbenchmark, BERT-base full-size timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python tools/gpu_fc_epi.py 2>&1 | grep "^K" | tee gpurun_out/q_fc_epi.log
timeout 400 python tools/gpu_configs.py bert > gpurun_out/q_bert_gelu.log 2>&1
grep -E "ms/|rel" gpurun_out/q_bert_gelu.log | head; grep -E "^ +[0-9.]+ us +[0-9.]+% x" gpurun_out/q_bert_gelu.log
