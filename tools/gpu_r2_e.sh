#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_pipeline.py tests/test_gpu_png.py tests/test_gpu_processor_callthrough.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_pre.log 2>&1; echo "pytest preprocess exit $?"; tail -8 gpurun_out/pytest_pre.log | cut -c1-300
for tw in 84 56 112; do
KOCR_PRE_TW=$tw python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-library-baselines --no-c5 > gpurun_out/bench_tw$tw.json 2> gpurun_out/bench_tw$tw.err; echo "bench tw=$tw exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_tw$tw.json')); print('tw $tw', d['value'], d['kernels']['preprocess'], d['preprocess_hbm'])"
done
python bench.py --workload c4 --steps 3 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4.json')); print('c4', d['value'], d['kernels']['preprocess'], d['preprocess_hbm'])"
python tools/prof_target.py 64 > gpurun_out/prof_plain.log 2>&1 && \
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"preprocess_kernel" -s 1 -c 1 -f -o gpurun_out/pre_r2 python tools/prof_target.py 64 > gpurun_out/prof_ncu.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/pre_r2.ncu-rep --page raw --csv > gpurun_out/pre_r2_raw.csv 2>/dev/null
ncu -i gpurun_out/pre_r2.ncu-rep --page source --csv > gpurun_out/pre_r2_source.csv 2>/dev/null
timeout 1200 python -m pytest tests/test_gpu_vllm_plugin.py -q -m gpu -p no:cacheprovider -s -x -rs > gpurun_out/pytest_vllm_plugin.log 2>&1; echo "pytest vllm plugin exit $?"; grep -v "^\[rank0\]\|Warning\|warn" gpurun_out/pytest_vllm_plugin.log | tail -30 | cut -c1-1200
