#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "attention" > gpurun_out/pytest_attn.log 2>&1; echo "pytest attention exit $?"; tail -6 gpurun_out/pytest_attn.log | cut -c1-300
KOCR_LIB=$PWD/karanta_ocr_b200/libkocr.so timeout 300 python tools/attn_bench.py 64 2>&1 | tail -1 | tee gpurun_out/attn_variants.txt
KOCR_LIB=$PWD/karanta_ocr_b200/libkocr_trace.so timeout 300 python tools/attn_trace.py > gpurun_out/attn_trace.txt 2>&1; echo "trace exit $?"; head -6 gpurun_out/attn_trace.txt | cut -c1-420
timeout 900 python -m pytest tests/test_gpu_tower.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_tower.log 2>&1; echo "pytest tower exit $?"; tail -3 gpurun_out/pytest_tower.log | cut -c1-300
