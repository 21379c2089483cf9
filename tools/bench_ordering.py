"""Timing of the ordering kernels (K8) on a device-resident discrepancy matrix (GPU box):
N instances -> all-vs-all (K6) -> tree penalty -> both treePenalizedPathLength copies; the oracle (the reference's
algorithm, pinned) on the host beside it for the smaller sizes."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from conftest import load_golden
from fr3d_python_b200 import synthetic
from fr3d_python_b200.geometry import discrepancy as disc
from fr3d_python_b200.ordering import orderBySimilarity as obs
from fr3d_python_b200.search import orderBySimilarity as sobs
from oracle import ordering_oracle as oo

av = load_golden("geometry.json.gz")["all_vs_all"]
sizes = [int(v) for v in sys.argv[1:]] or [300, 1000, 3000]


def timed(fn, *a):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(*a); torch.cuda.synchronize()
    return r, 1e3 * (time.perf_counter() - t0)


for N in sizes:
    C, R = synthetic.make_instances(av["base_centers"], av["base_rotations"], N, 5000)
    dm = disc.all_vs_all(C, R, out_device=True)
    obs.treePenalty(dm)                                   # warm-up (allocator, module load)
    _, t_pen = timed(obs.treePenalty, dm)
    o1, t_search = timed(sobs.treePenalizedPathLength, dm, 100, 59)
    o2, t_tppl = timed(obs.treePenalizedPathLength, dm, 10, 59)
    _, t_reorder = timed(obs.reorderSymmetricMatrix, dm, o2)
    line = {"n": N, "tree_penalty_ms": t_pen, "search_copy_100_reps_ms": t_search, "ordering_copy_10_reps_2opt_ms": t_tppl,
            "reorder_ms": t_reorder}
    if N <= 1000:
        host = dm.cpu().numpy()
        t0 = time.perf_counter(); r1 = oo.treePenalizedPathLength_search(host, 100, 59); line["oracle_search_copy_ms"] = 1e3 * (time.perf_counter() - t0)
        t0 = time.perf_counter(); r2 = oo.treePenalizedPathLength(host, 10, 59); line["oracle_ordering_copy_ms"] = 1e3 * (time.perf_counter() - t0)
        line["equal_to_oracle"] = bool(r1 == o1 and r2 == o2)
    print(json.dumps(line), flush=True)
