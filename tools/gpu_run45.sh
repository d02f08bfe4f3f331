#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload c5 --job-pages 2048 --steps 1 --warmup 3 > gpurun_out/bench_c5_2048.json 2> gpurun_out/bench_c5.err
echo "rc $?"; cat gpurun_out/bench_c5_2048.json | cut -c1-900; tail -3 gpurun_out/bench_c5.err
