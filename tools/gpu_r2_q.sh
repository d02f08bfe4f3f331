#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_png.py tests/test_gpu_pipeline.py tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_png.log 2>&1; echo "pytest png/pipeline/kernels exit $?"; tail -3 gpurun_out/pytest_png.log | cut -c1-300
timeout 600 python tools/png_overlap_bench.py 8 2>&1 | grep -v Warning | tail -8 | tee gpurun_out/png_overlap.txt
python bench.py --workload c5 --job-pages 512 --png --steps 1 --warmup 1 > gpurun_out/bench_c5_png.json 2> gpurun_out/bench_c5_png.err; echo "c5 png exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c5_png.json')); print('c5 png', d['value'], d.get('ms_per_step'))"
