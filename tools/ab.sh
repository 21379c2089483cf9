#!/bin/bash
# A/B on the GPU box: warps per SM of the single-precision stacking kernel (separate builds, -DFR3D_ST32_WARPS=n)
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_st_16.log 2>&1
for v in 12 20 24; do
  FR3D_B200_LIB=$PWD/fr3d_python_b200/lib/libfr3d_b200_st$v.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_st_$v.log 2>&1
done
