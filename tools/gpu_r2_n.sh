#!/bin/bash
# Round 2 pass N: preprocess kernel (incremental addressing in the horizontal pass, cheaper LUT addressing): parity + timing.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_pipeline.py tests/test_gpu_png.py tests/test_gpu_processor_callthrough.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_pre.log 2>&1; echo "pytest preprocess exit $?"; tail -3 gpurun_out/pytest_pre.log | cut -c1-300
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-library-baselines --no-c5 > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n.json')); print('c2', d['value'], d['kernels']['preprocess'], d['preprocess_hbm'])"
python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-library-baselines --no-c5 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4.json')); print('c4', d['value'], d['kernels']['preprocess'], d['preprocess_hbm'])"
python tools/prof_target.py 64 > gpurun_out/prof_plain.log 2>&1 && \
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"preprocess_kernel" -s 1 -c 1 -f -o gpurun_out/pre_r2c python tools/prof_target.py 64 > gpurun_out/prof_ncu.log 2>&1
echo "ncu preprocess exit $?"
ncu -i gpurun_out/pre_r2c.ncu-rep --page raw --csv > gpurun_out/pre_r2c_raw.csv 2>/dev/null
ncu -i gpurun_out/pre_r2c.ncu-rep --page source --csv > gpurun_out/pre_r2c_source.csv 2>/dev/null
