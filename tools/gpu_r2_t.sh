#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/graph_latency.py 2>&1 | grep -v Warning | tail -4 | tee gpurun_out/graph_latency.txt
for p in 1 4 16; do python bench.py --pages $p --steps 10 --warmup 3 --no-cpu-baseline --no-library-baselines --no-c5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'pages': d['config']['pages_per_step_per_gpu'], 'ms_per_step': d['ms_per_step'], 'pages_per_s': d['value'], 'e2e_pages_per_s': d['e2e']['value'], 'e2e_ms_per_step': d['e2e']['ms_per_step']}))"; done | tee gpurun_out/small_batches.jsonl
