#!/bin/bash
# exact-width base-2 softmax: parity tests, BERT-base full-size timing, ncu of the kernel alone (scores [16,12,384,384])
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_membound.py tests/test_gpu_golden.py tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -5
timeout 400 python tools/gpu_configs.py bert > gpurun_out/q_bert_exact.log 2>&1
grep -E "ms/|rel|Softmax" gpurun_out/q_bert_exact.log | head -30
sed -n '/^cat > \/tmp\/sm_run.py/,/^PY$/p' tools/gpu_call_r.sh | sed '1d;$d' > /tmp/sm_run.py
for ex in 1 0; do
B200_SOFTMAX_EXACT=$ex B200_NO_CUDA_GRAPH=1 timeout 300 ncu --set full --clock-control none -k regex:"softmax_" -s 1 -c 1 -o gpurun_out/q_softmax_exact$ex -f python /tmp/sm_run.py > gpurun_out/q_softmax_exact$ex.log 2>&1
tail -1 gpurun_out/q_softmax_exact$ex.log
done
