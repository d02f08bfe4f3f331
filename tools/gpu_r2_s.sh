#!/bin/bash
mkdir -p gpurun_out
python tools/png_prof_target.py > gpurun_out/png_plain.log 2>&1 && \
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"inflate_kernel" -s 1 -c 1 -f -o gpurun_out/inflate python tools/png_prof_target.py > gpurun_out/png_ncu.log 2>&1
echo "ncu inflate exit $?"
ncu -i gpurun_out/inflate.ncu-rep --page raw --csv > gpurun_out/inflate_raw.csv 2>/dev/null
ncu -i gpurun_out/inflate.ncu-rep --page source --csv > gpurun_out/inflate_source.csv 2>/dev/null
