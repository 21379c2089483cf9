#!/bin/bash
# A/B on the GPU box: min blocks per SM of the two thread-per-pair list kernels (separate builds)
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_occ_default.log 2>&1
for v in LIGHT_MINB_6 LIGHT_MINB_8 BPH_MINB_5 BPH_MINB_7; do
  FR3D_B200_LIB=$PWD/fr3d_python_b200/lib/libfr3d_b200_$v.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_occ_$v.log 2>&1
done
