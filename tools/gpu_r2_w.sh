#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "attention" > gpurun_out/pytest_attn3.log 2>&1; echo "pytest attention exit $?"; tail -2 gpurun_out/pytest_attn3.log | cut -c1-200
timeout 300 python tools/attn_bench.py 64 2>&1 | tail -1 | tee gpurun_out/attn3b.txt
KOCR_ATTN2=1 timeout 300 python tools/attn_bench.py 64 2>&1 | tail -1 | tee -a gpurun_out/attn3b.txt
KOCR_LIB=$PWD/karanta_ocr_b200/libkocr_trace.so timeout 300 python tools/attn_trace.py > gpurun_out/attn3_trace.txt 2>&1; head -3 gpurun_out/attn3_trace.txt | cut -c1-420; grep "^warp [123]:" gpurun_out/attn3_trace.txt
