"""Soak test of the single-precision kernels: the hit stream of the default path against the all-fp64 path
(FR3D_SCREEN=64) on many perturbed structures.  Usage (GPU box):
    python tools/soak_fp32.py run <out.npy> <n_structures> <seed0> <sigma>      (one path, per FR3D_SCREEN)
    python tools/soak_fp32.py <n_structures> <seed0> <sigma>                    (both paths in subprocesses + compare)
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(out, n, seed0, sigma):
    from conftest import ALL9
    from fr3d_python_b200 import synthetic
    from fr3d_python_b200.classifiers import NA_pairwise_interactions as NA
    from fr3d_python_b200.engine import Engine
    batch = synthetic.make_structure_batch(synthetic.pack_base(), n, seed0=seed0, sigma=sigma)
    bt, _ = NA._box_table(NA._default_focused(ALL9['basepair']), NA.load_ideal_basepair_hydrogen_bonds())
    hits, n_pairs, ex = Engine.get().annotate_batch(batch, bt, NA.category_mask(ALL9))
    order = np.lexsort((hits['slot'], hits['j'], hits['i'], hits['rank']))
    np.save(out, hits[order])
    print(n_pairs, len(hits), ex['plan'].counters.cpu().numpy().tolist())


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]))
    else:
        n, seed0, sigma = sys.argv[1:4]
        outs = []
        for mode in ("32", "64"):
            out = "/tmp/soak_%s.npy" % mode
            res = subprocess.run([sys.executable, __file__, "run", out, n, seed0, sigma], check=True,
                                 env=dict(os.environ, FR3D_SCREEN=mode), capture_output=True, text=True)
            print("FR3D_SCREEN=%s:" % mode, res.stdout.strip().splitlines()[-1])
            outs.append(np.load(out))
        same = outs[0].tobytes() == outs[1].tobytes()
        print("n=%s seed0=%s sigma=%s: %d hits, byte-identical: %s" % (n, seed0, sigma, len(outs[0]), same))
        sys.exit(0 if same else 1)
