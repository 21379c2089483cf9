#!/bin/bash
# A/B on the GPU box: stacking / pairing lists with the single-phase kernels (0) or the two-phase kernels (1)
for m in 0 1; do
  FR3D_LIST2=$m python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_list2_$m.log 2>&1
done
