#!/bin/bash
# A/B on the GPU box: single-precision screen with 32 warps per SM (default build), with 24 (a second build,
# -DFR3D_S32_WARPS=24 -> lib/libfr3d_b200_w24.so), and the double-precision screen (FR3D_SCREEN=64)
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_s32_w32.log 2>&1
FR3D_B200_LIB=$PWD/fr3d_python_b200/lib/libfr3d_b200_w24.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_s32_w24.log 2>&1
FR3D_SCREEN=64 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_s64.log 2>&1
