#!/bin/bash
# Round-end style run: the driver's own commands (pytest -m gpu, smoke, bench both arms) + ncu evidence.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_final.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['tensor_pipe']['frac_of_sustained_peak'], d['clocks'], d.get('cpu_baseline'))
for k, v in d['kernels'].items(): print(k, round(v['ms_per_launch'], 3), v['launches'], round(v.get('tflops', 0), 1))
print(d['roofline']); print(d['preprocess_hbm'])
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference exit $?"; cut -c1-400 gpurun_out/bench_reference.json
# launch list of the bench command itself (hot-path kernels only; weight prepack launches filtered out)
python bench.py --steps 1 --warmup 1 --pages 8 --no-cpu-baseline > gpurun_out/plain_p8.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_kernel|attention_kernel|norm_kernel|preprocess_kernel|cast_f32|gather_groups" -c 1200 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --pages 8 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
# full counters, one launch of each hot kernel at the C2 batch size (64 pages, depth 1)
python tools/prof_target.py 64 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"attention_kernel|preprocess_kernel|gemm_kernel|norm_kernel" -c 24 -f -o gpurun_out/prof_final python tools/prof_target.py 64 > gpurun_out/prof_ncu.log 2>&1
echo "ncu full exit $?"
