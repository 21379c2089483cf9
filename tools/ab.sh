#!/bin/bash
# A/B of classify kernel variants on a GPU box: prints kernel_ms per variant
mkdir -p gpurun_out
for lib in ab_u1 ab_u2 ab_pm_mb5 ab_pm_mb6; do
  for mult in 5 16; do
    out=$(FR3D_B200_LIB=$PWD/fr3d_python_b200/lib/$lib.so FR3D_CLASSIFY_GRID_MULT=$mult timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['kernel_ms']['classify_pairs'], d['value'], d['e2e']['value'])")
    echo "$lib mult=$mult classify_ms,value,e2e: $out"
  done
done
