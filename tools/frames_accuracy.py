"""Max deviation of the K1 frames from the numpy SVD route (oracle) on perturbed 1GID structures and on random
rigid motions that include near-180-degree rotations.  Usage (GPU box): python tools/frames_accuracy.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fr3d_oracle as O  # noqa: E402
from fr3d_python_b200 import synthetic  # noqa: E402
from fr3d_python_b200.geometry import frames as gframes  # noqa: E402


def main():
    base = synthetic.pack_base()
    batch = synthetic.make_structure_batch(base, 12, seed0=4000)
    status = gframes.fit_batch_frames(batch, place_h=True)
    assert np.all(status == 0)
    worst_r = worst_c = worst_orth = 0.0
    for s in range(batch.n_structures):
        spec = synthetic.batch_to_spec(batch, s)
        nts = O.build_nts(spec)
        lo = int(batch.struct_off[s])
        for k, nt in enumerate(nts):
            R = batch.rot[lo + k].reshape(3, 3)
            worst_r = max(worst_r, float(np.abs(R - nt.rotation).max()))
            worst_c = max(worst_c, float(np.abs(batch.center[lo + k] - nt.base_center).max()))
            worst_orth = max(worst_orth, float(np.abs(R @ R.T - np.eye(3)).max()))
    print("frames vs numpy SVD on %d nts: max |dR| %.3e  max |dcentre| %.3e  max |R R^T - I| %.3e"
          % (batch.n, worst_r, worst_c, worst_orth))


if __name__ == "__main__":
    main()
