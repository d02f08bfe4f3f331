#!/bin/bash
# Round-end style run: the driver's own commands (pytest -m gpu, smoke, bench both arms) + ncu evidence.
mkdir -p gpurun_out
if [ "$1" != "--no-tests" ]; then
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
fi
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_final.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['tensor_pipe']['frac_of_sustained_peak'], d['clocks'], d.get('cpu_baseline'))
for k, v in d['kernels'].items(): print(k, round(v['ms_per_launch'], 3), v['launches'], round(v.get('tflops', 0), 1))
print(d['roofline']); print(d['preprocess_hbm'])
PY
[ "$2" == "--no-reference" ] || python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference exit $?"; cut -c1-400 gpurun_out/bench_reference.json
# launch list of the bench command itself (hot-path kernels only; weight prepack launches filtered out)
python bench.py --steps 1 --warmup 1 --pages 8 --no-cpu-baseline --no-library-baselines --no-c5 > gpurun_out/plain_p8.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_kernel|attention|norm_kernel|preprocess_kernel|cast_f32|gather_groups" -c 1200 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --pages 8 --no-cpu-baseline --no-library-baselines --no-c5 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
# full counters, one launch of each hot kernel at the C2 batch size (64 pages, depth 1)
python tools/prof_target.py 64 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:"attention|preprocess_kernel|gemm_kernel|^norm_kernel|kocr::norm_kernel" -s 10 -c 10 -f -o gpurun_out/prof_final python tools/prof_target.py 64 > gpurun_out/prof_ncu.log 2>&1
echo "ncu full exit $?"
# the raw page travels back as csv; the report itself only if it is small enough for the 64 MiB return limit
ncu -i gpurun_out/prof_final.ncu-rep --page raw --csv > gpurun_out/prof_final_raw.csv 2>/dev/null
[ $(stat -c %s gpurun_out/prof_final.ncu-rep) -gt 40000000 ] && rm -f gpurun_out/prof_final.ncu-rep
du -sh gpurun_out
