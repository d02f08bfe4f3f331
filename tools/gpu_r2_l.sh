#!/bin/bash
# Round 2 pass L: attention timing probes by instruction class (results are wrong by construction; timing only).
mkdir -p gpurun_out
for v in "" _p1_noexp _p2_nomax _p16_noacc _p32_nosub _p64_nocvt _p114_noacc_nosub_nocvt_nomax _p115_only_ldst; do
KOCR_LIB=$PWD/karanta_ocr_b200/libkocr$v.so timeout 300 python tools/attn_bench.py 64 2>&1 | tail -2 | cut -c1-200
done | tee gpurun_out/attn_probes2.txt
