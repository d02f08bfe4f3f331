#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tower.py tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_mixed.log 2>&1; echo "pytest kernels/tower/pipeline exit $?"; grep -E "passed|failed" gpurun_out/pytest_mixed.log | tail -1
timeout 300 python tools/attn_bench.py 64 2>&1 | tail -1 | tee gpurun_out/attn_skip.txt
KOCR_ATTN_MIXED=1 timeout 300 python tools/attn_bench.py 64 2>&1 | tail -1 | tee -a gpurun_out/attn_skip.txt
KOCR_ATTN2=1 timeout 300 python tools/attn_bench.py 64 2>&1 | tail -1 | tee -a gpurun_out/attn_skip.txt
timeout 300 python tools/attn_bench.py 64 2>&1 | tail -1 | tee -a gpurun_out/attn_skip.txt
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-library-baselines --no-c5 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_x.json')); print('c2', d['value'], d['e2e']['value'], d['kernels']['attention'], d['clocks'])"
python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-library-baselines --no-c5 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4.json')); print('c4', d['value'], d['kernels']['attention'])"
