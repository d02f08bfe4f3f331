#!/bin/bash
# one process per GPU under torchrun: the weak-scaling bench line and the C5 bulk job at N GPUs
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --workload c5 --job-pages 8192 --steps 1 --warmup 3 > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err; echo "c5 n$N rc $?"
python - <<PY
import json
for f in ["bench_n$N", "bench_c5_n$N"]:
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["value"], 1), round(d["e2e"]["value"], 1), d["ms_per_step"], d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/bench_n$N.err
