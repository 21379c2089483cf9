"""Secondary measurements (not the driver's bench line): the other BASELINE.json configs.

  python tools/bench_other.py annotate   C1 / C2-like / C3 single structures through the public API
  python tools/bench_other.py disc       C5 all-vs-all discrepancy, N instances x n positions
Prints one JSON object per measurement.  CPU columns time the oracle port on a bounded sample.
"""

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fr3d_python_b200 import synthetic  # noqa: E402
from fr3d_python_b200.classifiers import NA_pairwise_interactions as NA  # noqa: E402
from fr3d_python_b200.engine import Engine, pin_batch  # noqa: E402
from fr3d_python_b200.geometry import discrepancy as disc  # noqa: E402

ALL9 = {'basepair': NA.Leontis_Westhof_basepairs + ['cWB', 'cBW'], 'basepair_detail': [], 'stacking': [], 'sO': [],
        'backbone': [], 'coplanar': [], 'sugar_ribose': [], 'covalent': [], 'near': []}


def annotate_configs(cpu=True):
    import torch
    from oracle import fr3d_oracle as O
    eng = Engine.get()
    bt, atoms = NA._box_table(NA._default_focused(ALL9['basepair']), NA.load_ideal_basepair_hydrogen_bonds())
    mask = NA.category_mask(ALL9)
    base = synthetic.pack_base()
    cases = [("C1 1GID-full (316 nt)", base),
             ("C2 S_1S72_like (2 844 nt)", synthetic.make_tiled_structure(base, 9, (3, 3, 1), 1072, pdb="1S72L")),
             ("C3 25 280-nt assembly", synthetic.make_tiled_structure(base, 80, (5, 4, 4), 25000, pdb="C3"))]
    for name, batch in cases:
        pinned = pin_batch(torch, batch)
        hits, n_pairs, ex = eng.annotate_batch(batch, bt, mask, pinned=pinned)
        plan = ex["plan"]
        for _ in range(3):
            eng.annotate_batch(batch, bt, mask, pinned=pinned, plan=plan)
        torch.cuda.synchronize()
        reps = 20
        t0 = time.perf_counter()
        for _ in range(reps):
            eng.annotate_batch(batch, bt, mask, pinned=pinned, plan=plan)
        torch.cuda.synchronize()
        t_dev = (time.perf_counter() - t0) / reps
        NA.annotate_batch(batch, ALL9)                 # first call of the process pays one-time initialisation
        t0 = time.perf_counter()
        res = NA.annotate_batch(batch, ALL9)
        t_full = time.perf_counter() - t0
        out = {"config": name, "nt": batch.n, "pairs": n_pairs, "hits": len(hits),
               "records_in_to_hits_out_ms": 1e3 * t_dev, "pairs_per_s": n_pairs / t_dev,
               "with_host_tail_to_tuples_ms": 1e3 * t_full, "rows": sum(len(v) for v in res[0][0].values())}
        if cpu and batch.n <= 3000:
            spec = synthetic.batch_to_spec(batch, 0)
            t0 = time.perf_counter()
            O.annotate(spec, ALL9)
            t_cpu = time.perf_counter() - t0
            out["cpu_oracle_1core_s"] = t_cpu
            out["cpu_pairs_per_s"] = n_pairs / t_cpu
        print(json.dumps(out))


def disc_configs():
    import gzip
    import torch
    from oracle import fr3d_oracle as O
    with gzip.open(os.path.join(ROOT, "tests", "golden", "geometry.json.gz"), "rt") as f:
        av = json.load(f)["all_vs_all"]
    eng = Engine.get()
    for N in (2000, 20000):
        C, R = synthetic.make_instances(av["base_centers"], av["base_rotations"], N, 5000)
        n = C.shape[1]
        dc = torch.from_numpy(C).to(eng.device)
        dr = torch.from_numpy(R.reshape(N, n, 9)).to(eng.device)
        out = torch.empty((N, N), dtype=torch.float64, device=eng.device)
        from fr3d_python_b200 import _lib

        def launch():
            _lib.check(eng.lib.fr3d_discrepancy_all_vs_all(dc.data_ptr(), dr.data_ptr(), None, N, n, 1.0, 0, N,
                                                           out.data_ptr(), eng.stream_ptr()), "all_vs_all")
        for _ in range(2):
            launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            launch()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        pairs = N * (N - 1) // 2         # evaluated once per unordered pair, mirrored into both triangles
        flop = pairs * (154 * n + 600)
        t0 = time.perf_counter()
        D = disc.all_vs_all(C, R)
        t_api = time.perf_counter() - t0
        res = {"config": "C5 all-vs-all discrepancy", "N": N, "n": n, "kernel_ms": ms,
               "instance_pairs_per_s": pairs / (ms / 1e3), "model_gflops": flop / (ms / 1e3) / 1e9,
               "api_host_in_host_out_s": t_api, "out_bytes": N * N * 8}
        # CPU: a bounded block with the oracle port
        m = 60
        t0 = time.perf_counter()
        O.all_vs_all_discrepancy(C[:m], R[:m])
        t_cpu = time.perf_counter() - t0
        res["cpu_oracle_pairs_per_s_1core"] = (m * (m - 1) / 2) / t_cpu
        res["max_abs_diff_vs_oracle_block"] = float(np.abs(D[:m, :m] - O.all_vs_all_discrepancy(C[:m], R[:m])).max())
        print(json.dumps(res))
        del out, dc, dr, D


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "annotate"
    if what == "annotate":
        annotate_configs()
    else:
        disc_configs()
