#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_preprocess.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench17.json 2> gpurun_out/bench17.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench17.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['tensor_pipe']['frac_of_sustained_peak'], d['clocks'])
for k, v in d['kernels'].items(): print(k, round(v['ms_per_launch'], 3), v['launches'], round(v.get('tflops', 0), 1))
print(d['preprocess_hbm'])
PY
tail -3 gpurun_out/bench17.err
