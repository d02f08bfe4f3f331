#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_pre.log 2>&1; echo "pytest preprocess exit $?"; tail -3 gpurun_out/pytest_pre.log | cut -c1-300
for tw in 84 112 56; do KOCR_PRE_TW=$tw python tools/pre_bench.py 2>&1 | tail -2; done | tee gpurun_out/pre_bench.txt
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-library-baselines --no-c5 > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n.json')); print('c2', d['value'], d['kernels']['preprocess'], d['preprocess_hbm'])"
