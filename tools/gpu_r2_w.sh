#!/bin/bash
mkdir -p gpurun_out
for v in "" _a3_tree _a3_early _a3_both _a3_both_poly8 ""; do
KOCR_LIB=$PWD/karanta_ocr_b200/libkocr$v.so timeout 300 python tools/attn_bench.py 64 2>&1 | tail -1 | tee -a gpurun_out/attn3_tune.txt
done
