#!/bin/bash
for rep in 1 2; do
for v in "" _at26 _at22 _at18; do
  for p in 16 64; do KOCR_LIB=$PWD/karanta_ocr_b200/libkocr$v.so timeout 300 python tools/attn_bench.py $p; done
done
done
