"""fr3d/geometry/RMSD.py:4-29 — trivial host reductions kept for API completeness (they are not on
the hot path: the kernels compute sse/rmsd themselves)."""

import numpy


def RMSD(set1, set2):
    assert len(set1) == len(set2)
    L = len(set1)
    assert L > 0
    return numpy.sqrt(numpy.sum(numpy.power(numpy.asarray(set1) - numpy.asarray(set2), 2)) / L)


def sumsquarederror(set1, set2):
    assert len(set1) == len(set2)
    assert len(set1) > 0
    return numpy.sum(numpy.power(numpy.asarray(set1) - numpy.asarray(set2), 2))
