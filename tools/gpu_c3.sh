#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 rc $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_c3.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"], {k:(round(x["ms_per_launch"],3), x["launches"]) for k,x in d["kernels"].items()})
PY
