#!/bin/bash
# A/B on the GPU box: pair_search occupancy variants (SEARCH_CAP candidates per staged segment, min blocks per SM),
# each a separate build: lib/libfr3d_b200_s<cap>_<minb>.so
for v in 64_10 96_9 64_12 128_7; do
  FR3D_B200_LIB=$PWD/fr3d_python_b200/lib/libfr3d_b200_s$v.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_search_$v.log 2>&1
done
