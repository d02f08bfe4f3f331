#!/bin/bash
# First-contact GPU run: each group in its own process (a faulting kernel poisons only its own CUDA context).
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
run() { # name, timeout, pytest args...
  local name=$1; local to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/$name.log | tee -a gpurun_out/summary.txt
}
run preprocess 400 tests/test_gpu_preprocess.py
run gemm_ones 120 "tests/test_gpu_kernels.py::test_gemm_ones_exact"
run gemm_identity 120 "tests/test_gpu_kernels.py::test_gemm_identity_picks_columns"
run gemm_plain 200 "tests/test_gpu_kernels.py::test_gemm_plain_and_bias"
run gemm_epi 200 "tests/test_gpu_kernels.py::test_gemm_epilogues" "tests/test_gpu_kernels.py::test_gemm_rejects_bad_shapes"
run norm 120 "tests/test_gpu_kernels.py::test_norm"
run attn_uniform 120 "tests/test_gpu_kernels.py::test_attention_uniform_probabilities"
run attn_order 120 "tests/test_gpu_kernels.py::test_attention_key_order"
run attn_random 300 "tests/test_gpu_kernels.py::test_attention_random"
run tower 900 tests/test_gpu_tower.py
