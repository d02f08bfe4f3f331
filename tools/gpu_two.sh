#!/bin/bash
# two GPUs of one box: the two-devices-in-one-process test and the 2-rank bench (weak scaling + extras)
mkdir -p gpurun_out
timeout 600 python -m pytest "tests/test_gpu_tower.py::test_two_devices_in_one_process" -q -m gpu -p no:cacheprovider > gpurun_out/pytest_two.log 2>&1; echo "two-device test exit $?"; tail -2 gpurun_out/pytest_two.log | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-c5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); print('n2', d['value'], d['e2e']['value'], d['clocks'])"
