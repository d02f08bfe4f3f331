#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k attention 2>&1 | tail -2
for rep in 1 2; do
for v in "" _nopf; do
  for p in 16 64; do KOCR_LIB=$PWD/karanta_ocr_b200/libkocr$v.so timeout 300 python tools/attn_bench.py $p; done
done
done
