#!/bin/bash
mkdir -p gpurun_out
python tools/prof_target.py 8 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"attention_kernel|preprocess_kernel|gemm_kernel" -s 9 -c 9 -f -o gpurun_out/prof_r1 python tools/prof_target.py 8 > gpurun_out/prof_ncu.log 2>&1
echo "ncu exit $?"; tail -5 gpurun_out/prof_ncu.log; ls -la gpurun_out/*.ncu-rep
