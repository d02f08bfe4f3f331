#!/bin/bash
mkdir -p gpurun_out
KOCR_LIB=$PWD/karanta_ocr_b200/libkocr_cs2.so python tools/prof_target.py 8 > gpurun_out/prof_plain.log 2>&1 && \
KOCR_LIB=$PWD/karanta_ocr_b200/libkocr_cs2.so ncu --set full --clock-control none --import-source on -k regex:"attention_kernel" -s 1 -c 1 -f -o gpurun_out/prof_attn_cs2 python tools/prof_target.py 8 > gpurun_out/prof_ncu.log 2>&1
echo "ncu exit $?"
