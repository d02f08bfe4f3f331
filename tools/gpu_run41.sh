#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
for v in "" _qsmem ""  _qsmem; do
  KOCR_LIB=$PWD/karanta_ocr_b200/libkocr$v.so timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench41$v.json 2> gpurun_out/bench41$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench41$v.json").read().strip().splitlines()[-1])
print("$v", d["value"], d["e2e"]["value"], d["clocks"], {k:round(x["ms_total"],1) for k,x in d["kernels"].items()})
PY
done
