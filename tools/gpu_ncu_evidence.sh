#!/bin/bash
# ncu evidence for the final kernels: launch list of the bench command (8 pages) and full counters at the C2 batch
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --pages 8 --no-cpu-baseline --no-library-baselines --no-c5 > gpurun_out/plain_p8.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_kernel|attention|norm_kernel|preprocess_kernel|cast_f32|gather_groups" -c 1200 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --pages 8 --no-cpu-baseline --no-library-baselines --no-c5 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
python tools/prof_target.py 64 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:"attention|preprocess_kernel|gemm_kernel|^norm_kernel|kocr::norm_kernel" -s 10 -c 10 -f -o gpurun_out/prof_final python tools/prof_target.py 64 > gpurun_out/prof_ncu.log 2>&1
echo "ncu full exit $?"
ncu -i gpurun_out/prof_final.ncu-rep --page raw --csv > gpurun_out/prof_final_raw.csv 2>/dev/null
rm -f gpurun_out/prof_final.ncu-rep
