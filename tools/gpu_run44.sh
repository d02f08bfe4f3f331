#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/hf_gpu_baseline.py 16 3 > gpurun_out/hf_gpu_baseline.jsonl 2> gpurun_out/hf_gpu_baseline.err
echo "rc $?"; cat gpurun_out/hf_gpu_baseline.jsonl; tail -5 gpurun_out/hf_gpu_baseline.err
