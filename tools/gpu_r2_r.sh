#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_png.py tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_png.log 2>&1; echo "pytest png/pipeline exit $?"; tail -3 gpurun_out/pytest_png.log | cut -c1-300
timeout 600 python tools/png_overlap_bench.py 8 2>&1 | grep -v Warning | tail -8 | tee gpurun_out/png_overlap.txt
