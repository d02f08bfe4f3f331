#!/bin/bash
# Round 2 pass M: persistent attention CTAs (both shapes): kernel + tower parity, timing (full shape isolated, C3 bench for the windowed one).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -x -k "attention" > gpurun_out/pytest_attn.log 2>&1; echo "pytest attention exit $?"; tail -4 gpurun_out/pytest_attn.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_tower.py tests/test_gpu_vllm_adapter.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_tower.log 2>&1; echo "pytest tower exit $?"; tail -4 gpurun_out/pytest_tower.log | cut -c1-300
for v in "" _pin; do
KOCR_LIB=$PWD/karanta_ocr_b200/libkocr$v.so timeout 300 python tools/attn_bench.py 64 2>&1 | tail -1
done | tee gpurun_out/attn_variants.txt
python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-library-baselines --no-c5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c3.json')); print('c3', d['value'], json.dumps(d['kernels']['attention']))"
python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-library-baselines --no-c5 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4.json')); print('c4', d['value'], json.dumps(d['kernels']['attention']))"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-library-baselines --no-c5 > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err; python -c "
import json; d=json.load(open('gpurun_out/bench_m.json')); print('c2', d['value'], json.dumps(d['kernels']['attention']))"
