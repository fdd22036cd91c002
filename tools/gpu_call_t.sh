#!/bin/bash
# final check of the round: FC epilogue microbenchmark + all four secondary configs, then the full round script (tests, bench, reference arm, ncu)
mkdir -p gpurun_out
python tools/gpu_fc_epi.py 2>&1 | grep "^K" | tee gpurun_out/t_fc_epi.log
timeout 600 python tools/gpu_configs.py > gpurun_out/t_configs.log 2>&1
grep -E "ms/|rel" gpurun_out/t_configs.log
bash tools/gpu_round.sh r18
