"""fr3d/geometry/angleofrotation.py:4-7.  A single 3x3 trace is host arithmetic; the batched angle
computation lives inside the discrepancy kernels (csrc/geometry.cuh: angle_from_trace)."""

import numpy as np


def angle_of_rotation(rotation_matrix):
    value = (np.trace(rotation_matrix) - 1.0) / 2.0
    value = np.clip(value, -1, 1)
    return np.arccos(value)
