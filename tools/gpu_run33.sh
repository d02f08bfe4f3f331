#!/bin/bash
for v in _p0 _p8 _p4 _p2; do
  for p in 16 64; do KOCR_LIB=$PWD/karanta_ocr_b200/libkocr$v.so timeout 300 python tools/attn_bench.py $p; done
done
