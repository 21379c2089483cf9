#!/bin/bash
# A/B of library builds on the GPU box: tools/ab_libs.sh <tag> lib1.so lib2.so ...  -> gpurun_out/ab_<tag>.log
# (each variant: bench.py resident + e2e figures, 20 steps)
tag=$1; shift
out=gpurun_out/ab_$tag.log
: > $out
for lib in "$@"; do
  echo "== $lib" >> $out
  FR3D_B200_LIB=$PWD/fr3d_python_b200/lib/$lib python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>>$out | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms_per_step %.4f' % d['ms_per_step'], d['kernel_ms'], 'e2e ms %.3f' % d['e2e']['ms_per_step'], 'pairs', d['pairs_per_step'], 'rows', d['rows_per_step'])" >> $out
done
cat $out
