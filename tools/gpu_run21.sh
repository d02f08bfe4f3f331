#!/bin/bash
mkdir -p gpurun_out
for w in c3 c4; do
python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w exit $?"
python - <<PY
import json
d = json.load(open('gpurun_out/bench_$w.json'))
print('$w value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['tensor_pipe']['frac_of_sustained_peak'], 'ms', d['ms_per_step'], d['clocks'])
for k, v in d['kernels'].items(): print('  ', k, round(v['ms_per_launch'], 3), v['launches'], round(v.get('tflops', 0), 1))
PY
tail -2 gpurun_out/bench_$w.err
done
