"""Save the hit records of a small synthetic batch (GPU box) so that the host tail can be profiled and
developed on a CPU-only machine: python tools/dump_hits.py gpurun_out/hits100.npy [n_structures]"""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import ALL9  # noqa: E402
from fr3d_python_b200 import synthetic  # noqa: E402
from fr3d_python_b200.classifiers import NA_pairwise_interactions as NA  # noqa: E402
from fr3d_python_b200.engine import Engine  # noqa: E402

n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
batch = synthetic.make_structure_batch(synthetic.pack_base(), n, seed0=4000)
bt, _ = NA._box_table(NA._default_focused(ALL9['basepair']), NA.load_ideal_basepair_hydrogen_bonds())
hits, n_pairs, _ = Engine.get().annotate_batch(batch, bt, NA.category_mask(ALL9))
np.save(sys.argv[1], hits)
print(len(hits), n_pairs)
