#!/bin/bash
# Round 2 pass H: preprocess kernel with the multi-column horizontal pass (parity + timing), TMEM rate micro, attention timing probes.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_pipeline.py tests/test_gpu_png.py tests/test_gpu_processor_callthrough.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_pre.log 2>&1; echo "pytest preprocess exit $?"; tail -8 gpurun_out/pytest_pre.log | cut -c1-300
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-library-baselines --no-c5 > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_h.json')); print('c2', d['value'], d['kernels']['preprocess'], d['preprocess_hbm'])"
python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-library-baselines --no-c5 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4.json')); print('c4', d['value'], d['kernels']['preprocess'], d['preprocess_hbm'])"
(cd tools/micro && timeout 120 ./tmem_rates) > gpurun_out/micro_tmem_rates.txt 2>&1; echo "tmem micro exit $?"; cat gpurun_out/micro_tmem_rates.txt
python tools/attn_bench.py 64 2>&1 | tail -1
for v in p1_noexp p2_nomax p3_noexp_nomax p4_noldtm p8_nosttm p15_shell pp0 poly0 poly2; do
KOCR_LIB=$PWD/karanta_ocr_b200/libkocr_$v.so timeout 300 python tools/attn_bench.py 64 2>&1 | tail -1
done | tee gpurun_out/attn_probes.txt
