"""One resident pass of the bench workload between cudaProfilerStart / Stop, for
   ncu --set full --profile-from-start off --clock-control none --import-source on -o gpurun_out/step_final python tools/one_pass.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import ALL9
from fr3d_python_b200 import synthetic, tail as tailmod
from fr3d_python_b200.classifiers import NA_pairwise_interactions as NA
from fr3d_python_b200.engine import Engine, DeviceBatch, DeviceIdent, pin_batch
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
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
torch.cuda.cudart().cudaProfilerStart()
eng.run_plan(db, plan, bt, mask, fit_frames=True, tail=(ident, labels))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("pairs", int(plan.counters[0]), "hits", int(plan.counters[1]), "rows", int(plan.tcounters[0]))
