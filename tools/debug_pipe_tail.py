"""Debug: pipeline + device tail vs single pass (GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import ALL9
from fr3d_python_b200 import synthetic
from fr3d_python_b200.classifiers import NA_pairwise_interactions as NA
from fr3d_python_b200.engine import Engine, Pipeline, pin_batch
base = synthetic.pack_base()
batch = synthetic.make_structure_batch(base, 37, seed0=900)
eng = Engine.get()
bt, atoms, labels = NA._tables(NA._default_focused(ALL9['basepair']), NA.load_ideal_basepair_hydrogen_bonds())
mask = NA.category_mask(ALL9)
ref = NA.annotate_batch(batch, ALL9, as_arrays=True)
print("ref counts", ref.row_count[:8], ref.n_rows)
pinned = pin_batch(torch, batch)
for chunks, caps in ((1, None), (5, None), (37, None), (3, (100, 100, 50))):
    kw = {} if caps is None else {"pair_cap": caps[0], "hit_cap": caps[1], "row_cap": caps[2]}
    pipe = Pipeline(eng, batch, pinned, bt, mask, n_chunks=chunks, labels=labels, **kw)
    for it in range(2):
        out = pipe.run()
        bad = []
        for s in range(batch.n_structures):
            got = out.rows[out.row_start[s]:out.row_start[s] + out.row_count[s]]
            if not np.array_equal(got, ref.rows_of(s)):
                bad.append(s)
        print("chunks", chunks, "caps", caps, "iter", it, "n_pairs", out.n_pairs, ref.n_pairs, "n_hits", out.n_hits, "rows", len(out.rows), "bad", bad[:10], len(bad))
        if bad:
            s = bad[0]
            print(" counts", out.row_count[:8], "starts", out.row_start[:8])
            print(" chunk plans", [(ch.plan.pair_cap, ch.plan.hit_cap, ch.plan.row_cap, ch.n_hits, ch.n_rows) for ch in pipe.chunks])
