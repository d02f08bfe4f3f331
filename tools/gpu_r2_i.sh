#!/bin/bash
# Round 2 pass I: attention CTA timeline (trace build), ncu source counters of the new preprocess kernel.
mkdir -p gpurun_out
KOCR_LIB=$PWD/karanta_ocr_b200/libkocr_trace.so timeout 300 python tools/attn_trace.py > gpurun_out/attn_trace.txt 2>&1; echo "trace exit $?"; head -14 gpurun_out/attn_trace.txt | cut -c1-420
python tools/prof_target.py 64 > gpurun_out/prof_plain.log 2>&1 && \
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"preprocess_kernel" -s 1 -c 1 -f -o gpurun_out/pre_r2b python tools/prof_target.py 64 > gpurun_out/prof_ncu.log 2>&1
echo "ncu preprocess exit $?"
ncu -i gpurun_out/pre_r2b.ncu-rep --page raw --csv > gpurun_out/pre_r2b_raw.csv 2>/dev/null
ncu -i gpurun_out/pre_r2b.ncu-rep --page source --csv > gpurun_out/pre_r2b_source.csv 2>/dev/null
