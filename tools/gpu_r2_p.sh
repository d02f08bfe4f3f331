#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/png_overlap_bench.py 8 2>&1 | grep -v Warning | tail -8 | tee gpurun_out/png_overlap.txt
