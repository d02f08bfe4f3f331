#!/bin/bash
# Round 2 pass G: TMEM load/store rates, ncu --set full with source counters for preprocess_kernel (C2 batch) and the windowed attention kernel (C3 widths).
mkdir -p gpurun_out
(cd tools/micro && timeout 120 ./tmem_rates) > gpurun_out/micro_tmem_rates.txt 2>&1; echo "tmem micro exit $?"; cat gpurun_out/micro_tmem_rates.txt
python tools/prof_target.py 64 > gpurun_out/prof_plain.log 2>&1 && \
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"preprocess_kernel" -s 1 -c 1 -f -o gpurun_out/pre_r2 python tools/prof_target.py 64 > gpurun_out/prof_ncu.log 2>&1
echo "ncu preprocess exit $?"
ncu -i gpurun_out/pre_r2.ncu-rep --page raw --csv > gpurun_out/pre_r2_raw.csv 2>/dev/null
ncu -i gpurun_out/pre_r2.ncu-rep --page source --csv > gpurun_out/pre_r2_source.csv 2>/dev/null
python tools/prof_target.py 64 qwen2_5_vl_7b > gpurun_out/prof_plain_q25.log 2>&1 && \
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"attention_kernel" -s 2 -c 2 -f -o gpurun_out/attn_q25 python tools/prof_target.py 64 qwen2_5_vl_7b > gpurun_out/prof_ncu_q25.log 2>&1
echo "ncu q25 attention exit $?"
ncu -i gpurun_out/attn_q25.ncu-rep --page raw --csv > gpurun_out/attn_q25_raw.csv 2>/dev/null
ncu -i gpurun_out/attn_q25.ncu-rep --page source --csv > gpurun_out/attn_q25_source.csv 2>/dev/null
ls -la gpurun_out; du -sh gpurun_out
