#!/bin/bash
# Round 2 GPU pass C: PNG decode parity, vLLM plugin inside a live engine.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_png.py -q -m gpu -p no:cacheprovider -s -x > gpurun_out/pytest_png.log 2>&1; echo "pytest png exit $?"; tail -15 gpurun_out/pytest_png.log
timeout 1600 python -m pytest tests/test_gpu_vllm_plugin.py -q -m gpu -p no:cacheprovider -s -x > gpurun_out/pytest_vllm_plugin.log 2>&1; echo "pytest vllm plugin exit $?"; tail -40 gpurun_out/pytest_vllm_plugin.log | cut -c1-600
