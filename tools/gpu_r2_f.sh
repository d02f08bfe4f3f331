#!/bin/bash
# Round 2 re-entry: establish state. Full parity suite, smoke, default bench, preprocess ncu.
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; grep -v "^\[rank0\]\|Warning\|warn" gpurun_out/pytest_gpu.log | tail -8 | cut -c1-400
grep -h "^parity\|^vllm adapter" gpurun_out/pytest_gpu.log > gpurun_out/parity_lines.txt; cat gpurun_out/parity_lines.txt | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; echo "bench exit $?"; cut -c1-3000 gpurun_out/bench_f.json
python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-library-baselines --no-c5 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4.json')); print('c4', d['value'], d['kernels']['preprocess'], d['preprocess_hbm'])"
python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-library-baselines --no-c5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c3.json')); print('c3', d['value'], json.dumps(d['kernels']))"
