#!/usr/bin/env bash
# quick GPU call: bench variants only (diagnostics)
set -x
mkdir -p gpurun_out
python bench.py --no-cpu-baseline > gpurun_out/q_bench_a.json 2>gpurun_out/q_a.err; cat gpurun_out/q_bench_a.json | cut -c1-120; python -c "
import json; d=json.load(open('gpurun_out/q_bench_a.json')); print('A value',d['value'],'bursts',d['burst_ms_per_step'],d['clocks'])"
python bench.py --no-cpu-baseline --no-clock-sampler > gpurun_out/q_bench_b.json 2>gpurun_out/q_b.err; python -c "
import json; d=json.load(open('gpurun_out/q_bench_b.json')); print('B(no sampler) value',d['value'],'bursts',d['burst_ms_per_step'])"
python bench.py --no-cpu-baseline --steps 200 > gpurun_out/q_bench_c.json 2>gpurun_out/q_c.err; python -c "
import json; d=json.load(open('gpurun_out/q_bench_c.json')); print('C(200 steps) value',d['value'],'bursts',d['burst_ms_per_step'],d['clocks'])"
