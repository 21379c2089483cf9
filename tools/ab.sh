#!/bin/bash
# A/B on the GPU box: number of chunks of the e2e pipeline with one host -> device copy per chunk
for c in 10 12 16 24; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --chunks $c > gpurun_out/ab_chunks_$c.log 2>&1
done
