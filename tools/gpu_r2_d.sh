#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_pipeline.py tests/test_gpu_png.py tests/test_gpu_processor_callthrough.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_pre.log 2>&1; echo "pytest preprocess exit $?"; tail -30 gpurun_out/pytest_pre.log | cut -c1-300
for tw in 84 112 56; do
KOCR_PRE_TW=$tw python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-library-baselines > gpurun_out/bench_tw$tw.json 2> gpurun_out/bench_tw$tw.err; echo "bench tw=$tw exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_tw$tw.json')); print('tw $tw', d['value'], d['kernels']['preprocess'], d['preprocess_hbm'])"
done
KOCR_PRE_TW=84 python bench.py --workload c4 --steps 3 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4.json')); print('c4', d['value'], d['kernels']['preprocess'], d['preprocess_hbm'])"
timeout 900 python -m pytest "tests/test_gpu_vllm_plugin.py::test_plugin_inside_live_engine[qwen2_vl]" -q -m gpu -p no:cacheprovider -s -x -rs > gpurun_out/pytest_vllm_plugin.log 2>&1; echo "pytest vllm plugin exit $?"; tail -60 gpurun_out/pytest_vllm_plugin.log | cut -c1-1500
