#!/usr/bin/env bash
# One gpurun call: parity tests, bench (+ reference arm), ncu launch list of the bench command, ncu traffic / tensor-pipe pass over every
# tcgen05 launch of one step, ncu full capture (with source) of a few pair-kernel launches.
# usage: bash tools/gpu_round.sh <tag>
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -8 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; cat gpurun_out/${TAG}_bench_ref.json; tail -3 gpurun_out/${TAG}_bench_ref.err
export B200_TUNE_FILE=$PWD/gpurun_out/${TAG}_tune.txt; rm -f $B200_TUNE_FILE; timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-latency > /dev/null 2>&1   # writes the kernel-selection file the ncu passes reuse (no trial launches under the profiler)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-latency > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum \
    --clock-control none -k regex:conv_tc -c 400 --csv --log-file gpurun_out/${TAG}_conv_traffic.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-latency > gpurun_out/${TAG}_ncu_traffic.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_tc3 -s 28 -c 6 -o gpurun_out/${TAG}_prof \
    python tools/gpu_r50_ab.py --one ncu > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
rm -f gpurun_out/ab_ncu.npy
ls -la gpurun_out | tail -12
