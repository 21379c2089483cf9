#!/bin/bash
# A/B on the GPU box: screen rows staged in 16-byte pieces and read with 128-bit loads (default build, 32 warps/SM;
# _w24: 24 warps/SM) against 4-byte copies and scalar loads (_scalar: -DFR3D_S32_SCALAR_ROWS)
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_rows_default.log 2>&1
for v in w24 scalar; do
  FR3D_B200_LIB=$PWD/fr3d_python_b200/lib/libfr3d_b200_$v.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_rows_$v.log 2>&1
done
