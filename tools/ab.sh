#!/bin/bash
# A/B on the GPU box: detail list read from global memory (0) or staged in shared memory (1)
for m in 0 1; do
  FR3D_DETAIL_STAGED=$m python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab_detail$m.log 2>&1
done
