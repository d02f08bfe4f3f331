#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { local name=$1; local to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/$name.log | tee -a gpurun_out/summary.txt
}
run preprocess 400 tests/test_gpu_preprocess.py
run kernels 600 tests/test_gpu_kernels.py
run tower 900 tests/test_gpu_tower.py
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench13.json 2> gpurun_out/bench13.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench13.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['tensor_pipe']['frac_of_sustained_peak'], d['clocks'])
for k, v in d['kernels'].items(): print(k, round(v['ms_per_launch'], 3), v['launches'], round(v.get('tflops', 0), 1))
PY
python tools/prof_target.py 8 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"preprocess_kernel" -s 1 -c 1 -f -o gpurun_out/prof_pp_v7 python tools/prof_target.py 8 > gpurun_out/prof_ncu.log 2>&1
echo "ncu exit $?"
bash tools/gpu_variants.sh
