"""Resident pass (records in HBM) as 1 / 2 / 4 / 8 chunks on three streams, one CUDA graph per pass: does overlapping one
chunk's tail with the next chunk's classify pay?  (GPU box)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import ALL9
from fr3d_python_b200 import parallel, synthetic
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
batch = synthetic.make_structure_batch(synthetic.pack_base(), S, seed0=4000)
for chunks in (1, 2, 3, 4, 6, 8):
    sh = parallel.ShardedAnnotator(batch, ALL9, rank=0, world_size=1, n_chunks=chunks)
    for _ in range(3):
        sh.run()
    for _ in range(3):
        sh.enqueue(upload=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        sh.enqueue(upload=False)
    e1.record()
    torch.cuda.synchronize()
    print("chunks %d: resident %.3f ms per pass" % (chunks, e0.elapsed_time(e1) / 20))
    del sh
