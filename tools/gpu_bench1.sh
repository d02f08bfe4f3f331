#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench exit $?"; tail -c 3000 gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
python bench.py --steps 1 --warmup 1 --pages 8 --no-cpu-baseline > gpurun_out/plain_p8.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --pages 8 --no-cpu-baseline > gpurun_out/ncu_p8.log 2>&1
echo "ncu exit $?"
