"""Where the e2e pass's time goes besides the kernels: the pipeline's own H2D blobs alone, with the D2H of the rows running
against them, and the whole pass (GPU box).  python tools/probe_pipeline_copies.py [chunks]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import ALL9
from fr3d_python_b200 import synthetic
from fr3d_python_b200.classifiers import NA_pairwise_interactions as NA
from fr3d_python_b200.engine import Engine, Pipeline, pin_batch
chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 8
batch = synthetic.make_structure_batch(synthetic.pack_base(), 1000, seed0=4000)
eng = Engine.get()
bt, atoms, labels = NA._tables(NA._default_focused(ALL9['basepair']), NA.load_ideal_basepair_hydrogen_bonds())
mask = NA.category_mask(ALL9)
pinned = pin_batch(torch, batch)
pipe = Pipeline(eng, batch, pinned, bt, mask, n_chunks=chunks, labels=labels)
for _ in range(3):
    pipe.run_graph()
torch.cuda.synchronize()


def timed(fn, reps=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d_one_stream():
    for ch in pipe.chunks:
        ch.dev_blob.copy_(ch.host_blob, non_blocking=True)


side = torch.cuda.Stream()
rows_dev = torch.empty(pipe.rows_host.numel(), dtype=torch.uint8, device=eng.device)


def h2d_with_d2h():
    with torch.cuda.stream(side):
        pipe.rows_host.copy_(rows_dev, non_blocking=True)
    h2d_one_stream()
    torch.cuda.current_stream().wait_stream(side)


total = sum(ch.host_blob.numel() for ch in pipe.chunks)
t = timed(h2d_one_stream)
print("H2D of the %d chunk blobs (%.1f MB) alone: %.3f ms = %.1f GB/s" % (chunks, total / 1e6, t, total / t / 1e6))
t2 = timed(h2d_with_d2h)
print("... with a %.1f MB D2H running against it: %.3f ms" % (pipe.rows_host.numel() / 1e6, t2))
print("whole pass (graph): %.3f ms" % timed(lambda: pipe.enqueue_graph(upload=True)))
print("resident pass (graph, no upload): %.3f ms" % timed(lambda: pipe.enqueue_graph(upload=False)))
