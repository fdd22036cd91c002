#!/usr/bin/env bash
# One gpurun call: parity tests, probe (per-kernel table), bench, ncu launch list, ncu full capture of the conv kernel.
# usage: bash tools/gpu_round.sh <tag> [ncu-kernel-regex] [skip] [count]
TAG=${1:-run}; KRE=${2:-conv_tc2}; SKIP=${3:-108}; CNT=${4:-8}
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -15 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python tools/gpu_probe.py > gpurun_out/${TAG}_probe.log 2>&1; grep -v "^  " gpurun_out/${TAG}_probe.log | tail -30
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; cat gpurun_out/${TAG}_bench_ref.json; tail -3 gpurun_out/${TAG}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-latency > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:${KRE} -s ${SKIP} -c ${CNT} -o gpurun_out/${TAG}_prof \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-latency > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum \
    --clock-control none -k regex:conv_tc2 -s 108 -c 54 --csv --log-file gpurun_out/${TAG}_conv_traffic.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-latency > gpurun_out/${TAG}_ncu_traffic.log 2>&1
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
ls -la gpurun_out
SZ=$(stat -c %s gpurun_out/${TAG}_prof.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -gt 45000000 ]; then rm -f gpurun_out/${TAG}_prof.ncu-rep; echo "dropped big ncu-rep ($SZ bytes), kept the raw csv"; fi
