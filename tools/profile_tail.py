"""Per-phase cycle counts of tail_structure_kernel (needs a library built with -DFR3D_TAIL_PROFILE:
   FR3D_B200_LIB=/path/to/profile.so python tools/profile_tail.py [structures [c2|c3]])."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import ALL9
from fr3d_python_b200 import synthetic, tail as tailmod
from fr3d_python_b200.classifiers import NA_pairwise_interactions as NA
from fr3d_python_b200.engine import Engine, DeviceBatch, DeviceIdent, pin_batch
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
if len(sys.argv) > 2:       # one tiled structure instead of the batch: c2 (2 844 nt) or c3 (25 280 nt)
    copies, lattice, seed = {"c2": (9, (3, 3, 1), 1072), "c3": (80, (5, 4, 4), 25000)}[sys.argv[2]]
    batch = synthetic.make_tiled_structure(synthetic.pack_base(), copies, lattice, seed=seed, pdb=sys.argv[2].upper())
    S = 1
else:
    batch = synthetic.make_structure_batch(synthetic.pack_base(), S, seed0=4000)
eng = Engine.get()
bt, atoms, labels = NA._tables(NA._default_focused(ALL9['basepair']), NA.load_ideal_basepair_hydrogen_bonds())
mask = NA.category_mask(ALL9)
pinned = pin_batch(torch, batch)
ident = DeviceIdent(torch, eng.device, tailmod.tail_ident(batch))
plan = eng.make_plan(batch.n, tail=(ident.n_struct, ident.n_ends))
db = DeviceBatch(torch, eng.device, batch, pinned)
for it in range(3):
    eng.run_plan(db, plan, bt, mask, fit_frames=True, tail=(ident, labels))
torch.cuda.synchronize()
plan.tcounters[8:].zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for it in range(reps):
    eng.run_tail(db, plan, (ident, labels))
e1.record()
torch.cuda.synchronize()
tc = plan.tcounters.cpu().numpy()
names = ["init", "sort hits", "walk", "entries: keys, sort, groups", "same-edge cleanup (1 thread)", "emission", "nested cWW",
         "crossing numbers", "label ranks + sort rows", "covalent sort + counts", "row allocation + writes"]
tot = tc[8:19].sum()
print("tail: %.3f ms per call (%d structures)" % (e0.elapsed_time(e1) / reps, S))
for k, nm in enumerate(names):
    print("  %-32s %10.0f cycles per structure  %5.1f %%" % (nm, tc[8 + k] / reps / S, 100.0 * tc[8 + k] / max(tot, 1)))
print("  total %.0f cycles per structure = %.1f us at 1.9 GHz" % (tot / reps / S, tot / reps / S / 1.9e3))
