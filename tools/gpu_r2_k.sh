#!/bin/bash
# Round 2 pass K: attention kernel with the row sum on the tensor pipe (ones column in V): kernel + tower parity, timing variants, timeline.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -x -k "attention" > gpurun_out/pytest_attn.log 2>&1; echo "pytest attention exit $?"; tail -4 gpurun_out/pytest_attn.log | cut -c1-300
for v in "" _poly0 _poly8 _poly3; do
KOCR_LIB=$PWD/karanta_ocr_b200/libkocr$v.so timeout 300 python tools/attn_bench.py 64 2>&1 | tail -1
done | tee gpurun_out/attn_variants.txt
KOCR_LIB=$PWD/karanta_ocr_b200/libkocr_trace.so timeout 300 python tools/attn_trace.py > gpurun_out/attn_trace.txt 2>&1; echo "trace exit $?"; head -14 gpurun_out/attn_trace.txt | cut -c1-420
timeout 900 python -m pytest tests/test_gpu_tower.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_tower.log 2>&1; echo "pytest tower exit $?"; tail -4 gpurun_out/pytest_tower.log | cut -c1-300
