#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tower.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
for v in "" _cs1; do
  KOCR_LIB=$PWD/karanta_ocr_b200/libkocr$v.so python tools/attn_bench.py 32 2>&1 | tail -1 | tee -a gpurun_out/attn_variants.txt
done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench15.json 2> gpurun_out/bench15.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench15.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['tensor_pipe']['frac_of_sustained_peak'], d['clocks'])
for k, v in d['kernels'].items(): print(k, round(v['ms_per_launch'], 3), v['launches'], round(v.get('tflops', 0), 1))
PY
tail -3 gpurun_out/bench15.err
