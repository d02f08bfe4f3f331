#!/bin/bash
# Round 2 GPU pass: parity suite, smoke, stall-reason capture of the attention kernel (+ preprocess kernel source counters), bench both arms.
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -5 gpurun_out/pytest_gpu.log
grep -h "^parity\|^vllm adapter" gpurun_out/pytest_gpu.log > gpurun_out/parity_lines.txt; cat gpurun_out/parity_lines.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
python tools/prof_target.py 64 > gpurun_out/prof_plain.log 2>&1 && \
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"attention_kernel|preprocess_kernel" -s 2 -c 2 -f -o gpurun_out/attn_r2 python tools/prof_target.py 64 > gpurun_out/prof_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/prof_ncu.log
python tools/ncu_stall_breakdown.py gpurun_out/attn_r2.ncu-rep > gpurun_out/attn_r2_stalls.txt 2>&1; echo "stall breakdown exit $?"; head -5 gpurun_out/attn_r2_stalls.txt
ncu -i gpurun_out/attn_r2.ncu-rep --page raw --csv > gpurun_out/attn_r2_raw.csv 2>/dev/null
ncu -i gpurun_out/attn_r2.ncu-rep --page source --csv > gpurun_out/attn_r2_source.csv 2>/dev/null
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench exit $?"; cut -c1-600 gpurun_out/bench_a.json
ls -la gpurun_out; du -sh gpurun_out
