"""Debug: repeated resident passes vs the calibration pass (GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import ALL9
from fr3d_python_b200 import synthetic, tail as tailmod
from fr3d_python_b200.classifiers import NA_pairwise_interactions as NA
from fr3d_python_b200.engine import Engine, DeviceBatch, DeviceIdent, pin_batch, HIT_DTYPE
S = int(sys.argv[1]) if len(sys.argv) > 1 else 200
base = synthetic.pack_base()
batch = synthetic.make_structure_batch(base, S, seed0=4000)
eng = Engine.get()
bt, atoms, labels = NA._tables(NA._default_focused(ALL9['basepair']), NA.load_ideal_basepair_hydrogen_bonds())
mask = NA.category_mask(ALL9)
pinned = pin_batch(torch, batch)
ident = DeviceIdent(torch, eng.device, tailmod.tail_ident(batch))
hits, n_pairs, ex = eng.annotate_batch(batch, bt, mask, pinned=pinned, labels=labels, want_hits=True)
print("calib", n_pairs, len(hits), np.bincount(hits['slot'], minlength=9))
for tight in (False, True):
    plan = eng.make_plan(batch.n, pair_cap=(n_pairs + 4096) if tight else None, hit_cap=(len(hits) + 4096) if tight else None,
                         tail=(ident.n_struct, ident.n_ends))
    db = DeviceBatch(torch, eng.device, batch, pinned)
    for it in range(3):
        eng.run_plan(db, plan, bt, mask, fit_frames=True, tail=(ident, labels))
        np_, nh = eng.read_counts(plan)
        h = plan.hits[:nh * 32].cpu().numpy().view(HIT_DTYPE)
        bx = db.t["base_xyz"].cpu().numpy()
        print("tight", tight, "it", it, np_, nh, np.bincount(h['slot'], minlength=9), "rows", eng.read_tail_counts(plan),
              "H slot sums", np.abs(bx[:, 11:]).sum(), "wl", plan.counters.cpu().tolist()[4:])
print("--- variants")
for name, kw_first, kw_next in (("no tail", dict(fit_frames=True), dict(fit_frames=True)),
                                ("tail, no refit", dict(fit_frames=True, tail=(ident, labels)), dict(fit_frames=False, tail=(ident, labels)))):
    plan = eng.make_plan(batch.n, tail=(ident.n_struct, ident.n_ends))
    db = DeviceBatch(torch, eng.device, batch, pinned)
    for it in range(3):
        m0 = db.t["base_mask"].clone() if it else None
        eng.run_plan(db, plan, bt, mask, **(kw_first if it == 0 else kw_next))
        np_, nh = eng.read_counts(plan)
        h = plan.hits[:nh * 32].cpu().numpy().view(HIT_DTYPE)
        bm = db.t["base_mask"].cpu().numpy().view(np.uint32)
        print(name, "it", it, np_, nh, np.bincount(h['slot'], minlength=9), "mask popcount", int(np.unpackbits(bm.view(np.uint8)).sum()),
              "bbmask sum", int(db.t["bb_mask"].cpu().numpy().view(np.uint32).astype(np.int64).sum()),
              "bb sum", float(db.t["bb_xyz"].abs().sum()), "meta sum", int(db.t["meta"].cpu().numpy().astype(np.int64).sum()))
