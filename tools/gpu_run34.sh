#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tower.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5
for p in 16 64; do timeout 300 python tools/attn_bench.py $p; done
