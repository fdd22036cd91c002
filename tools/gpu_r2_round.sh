#!/usr/bin/env bash
# Round-2 evidence pass on one GPU: ncu launch list of the bench command, ncu traffic / tensor-pipe pass over every tcgen05 launch of one
# step (roofline.traffic), bench + reference arm.  usage: bash tools/gpu_r2_round.sh <tag>
TAG=${1:-r2r}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json | cut -c1-400; tail -3 gpurun_out/${TAG}_bench.err
timeout 400 python bench.py --impl reference --steps 12 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; cat gpurun_out/${TAG}_bench_ref.json | cut -c1-300
export B200_TUNE_FILE=$PWD/gpurun_out/${TAG}_tune.txt; rm -f $B200_TUNE_FILE
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-latency --no-benchmark-app > /dev/null 2>&1   # writes the kernel-selection file the ncu passes reuse
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-latency --no-benchmark-app > gpurun_out/${TAG}_ncu_launches.log 2>&1
python tools/summarise_ncu.py launches gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.txt; cat gpurun_out/${TAG}_launches.txt
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum \
    --clock-control none -k regex:conv_tc -c 400 --csv --log-file gpurun_out/${TAG}_conv_traffic.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-latency --no-benchmark-app > gpurun_out/${TAG}_ncu_traffic.log 2>&1
python tools/summarise_ncu.py traffic gpurun_out/${TAG}_conv_traffic.csv > gpurun_out/${TAG}_conv_traffic.txt; tail -4 gpurun_out/${TAG}_conv_traffic.txt
ls -la gpurun_out | tail -8
