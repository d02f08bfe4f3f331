#!/bin/bash
# what the driver runs at round end, without the ncu passes: pytest -m gpu, smoke(), bench.py
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
python bench.py --no-cpu-baseline --no-library-baselines > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_check.json')); print(d['value'], d['e2e']['value'], d['tensor_pipe']['frac_of_sustained_peak'], d['gpu_launches'], d['clocks'])"
