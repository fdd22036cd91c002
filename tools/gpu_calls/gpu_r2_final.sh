#!/usr/bin/env bash
# final evidence pass of round 2 on one GPU: full GPU suite, smoke, bench + reference arm, ncu launch list + traffic pass, the three other configs
TAG=${1:-r2fin}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/ -m gpu -q -p no:cacheprovider --timeout 300 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python tools/gpu_configs.py mobilenet ssd bert > gpurun_out/${TAG}_configs.log 2>&1; grep -E "batch" gpurun_out/${TAG}_configs.log
bash tools/gpu_r2_round.sh ${TAG}
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_20steps.json 2> /dev/null; cut -c1-200 gpurun_out/${TAG}_bench_20steps.json
