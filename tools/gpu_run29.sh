#!/bin/bash
mkdir -p gpurun_out
python tools/prof_target.py 8 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"attention_kernel" -s 1 -c 1 -f -o gpurun_out/prof_attn_eo python tools/prof_target.py 8 > gpurun_out/prof_ncu.log 2>&1
echo "ncu exit $?"
