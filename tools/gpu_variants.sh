#!/bin/bash
mkdir -p gpurun_out
for v in "" _p3 _p5 _p6 _p8; do
  KOCR_LIB=$PWD/karanta_ocr_b200/libkocr$v.so python tools/attn_bench.py 32 2>&1 | tail -1 | tee -a gpurun_out/attn_variants.txt
done
