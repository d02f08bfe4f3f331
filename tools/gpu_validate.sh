#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -2
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['preprocess_hbm']['frac'], {k: round(v['value'],1) for k,v in d['extra'].items() if isinstance(v, dict) and 'value' in v})"
python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-library-baselines --no-c5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c3.json')); print('c3', d['value'], json.dumps(d['kernels'].get('attention_windowed')), json.dumps(d['kernels'].get('attention')))"
