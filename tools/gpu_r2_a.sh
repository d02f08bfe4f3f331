#!/bin/bash
# Round 2, first GPU pass: parity suite with the new depth-32 / call-through / plan-cache tests, instruction-rate micro
# benchmark (packed half-precision ex2), attention head-to-head vs FA4 / FA2, stall-reason capture of the attention kernel.
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -5 gpurun_out/pytest_gpu.log
grep -h "^parity\|^vllm adapter" gpurun_out/pytest_gpu.log > gpurun_out/parity_lines.txt; cat gpurun_out/parity_lines.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
(cd tools/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu_rates alu_rates.cu && timeout 120 ./alu_rates) > gpurun_out/micro_alu_rates.txt 2>&1; echo "micro exit $?"; grep -i "ex2\|f16x2\|bf16x2" gpurun_out/micro_alu_rates.txt
timeout 900 python tools/attn_head_to_head.py 64 10 > gpurun_out/attn_head_to_head.jsonl 2> gpurun_out/attn_head_to_head.err; echo "h2h exit $?"; cat gpurun_out/attn_head_to_head.jsonl | cut -c1-400
python tools/prof_target.py 64 > gpurun_out/prof_plain.log 2>&1 && \
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"attention_kernel" -s 1 -c 1 -f -o gpurun_out/attn_r2 python tools/prof_target.py 64 > gpurun_out/prof_ncu.log 2>&1
echo "ncu attention exit $?"
python tools/ncu_stall_breakdown.py gpurun_out/attn_r2.ncu-rep > gpurun_out/attn_r2_stalls.txt 2>&1; echo "stall breakdown exit $?"; head -5 gpurun_out/attn_r2_stalls.txt
ncu -i gpurun_out/attn_r2.ncu-rep --page raw --csv > gpurun_out/attn_r2_raw.csv 2>/dev/null
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench exit $?"; cut -c1-600 gpurun_out/bench_a.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_a.json 2> gpurun_out/bench_ref_a.err; echo "reference exit $?"; cut -c1-500 gpurun_out/bench_ref_a.json
ls -la gpurun_out; du -sh gpurun_out
