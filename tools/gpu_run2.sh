#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { local name=$1; local to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/$name.log | tee -a gpurun_out/summary.txt
}
run attn_uniform 120 "tests/test_gpu_kernels.py::test_attention_uniform_probabilities"
run attn_order 120 "tests/test_gpu_kernels.py::test_attention_key_order"
run attn_random 300 "tests/test_gpu_kernels.py::test_attention_random"
run tower 900 tests/test_gpu_tower.py
