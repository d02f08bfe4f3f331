#!/bin/bash
# Round 2 validation: the driver's commands + ncu evidence + C3/C4 lines.
bash tools/gpu_final.sh
python bench.py --workload c3 --steps 3 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c3.json')); print('c3', d['value'], json.dumps(d['kernels'].get('attention_windowed')), json.dumps(d['kernels'].get('attention')))"
python bench.py --workload c4 --steps 3 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c4.json')); print('c4', d['value'], d['preprocess_hbm'])"
python tools/prof_target.py 64 qwen2_5_vl_7b > gpurun_out/prof_plain_q25.log 2>&1 && \
timeout 600 ncu --set full --clock-control none -k regex:"attention_kernel" -s 2 -c 2 -f -o gpurun_out/attn_q25 python tools/prof_target.py 64 qwen2_5_vl_7b > gpurun_out/prof_ncu_q25.log 2>&1
ncu -i gpurun_out/attn_q25.ncu-rep --page raw --csv > gpurun_out/attn_q25_raw.csv 2>/dev/null; rm -f gpurun_out/attn_q25.ncu-rep
