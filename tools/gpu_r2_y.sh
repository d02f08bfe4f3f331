#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/attn_bench.py 64 > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:"attention" -s 4 -c 6 --csv --log-file gpurun_out/attn_mixed_launches.csv python tools/attn_bench.py 64 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/attn_mixed_launches.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size')
for r in rows[1:]:
    print(r[ki][:40], r[gi], r[mi], r[vi])
PY
