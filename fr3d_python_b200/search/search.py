"""FR3D search, the discrepancy screen of the possibilities: the loop of
fr3d/search/search.py:1067-1085 - for every possibility (a tuple of unit indices that met the pairwise constraints)
the geometric discrepancy to the query motif with early exit at the cutoff, keeping those below it - as ONE launch of
fr3d_discrepancy_one_vs_many over all possibilities instead of one matrix_discrepancy_cutoff call each.

The constraint intersection that produces the possibilities (search.py:263-375, a recursive backtracking search over
pair lists) and the orderings (fr3d/ordering) stay with the reference."""

import numpy as np

from ..geometry import discrepancy as disc


def possibility_discrepancies(querycenters, queryrotations, units, possibilities, cutoff, device=None):
    """search.py:1067-1085.  querycenters / queryrotations: the (permuted) query, one entry per position; units: the
    structure's units, units[i]["centers"] (3,), units[i]["rotations"] (3, 3) or an empty / None rotation; possibilities:
    iterable of index tuples.  Returns {possibility: discrepancy} for the possibilities with d is not None and
    d < cutoff, in the order of `possibilities` (the reference's dict insertion order)."""
    possibilities = [tuple(p) for p in possibilities]
    if not possibilities:
        return {}
    n = len(querycenters)
    idx = np.asarray(possibilities, dtype=np.int64)
    assert idx.shape[1] == n
    used = np.unique(idx)
    cen = np.zeros((int(used.max()) + 1, 3))
    rot = np.zeros((int(used.max()) + 1, 9))
    has = np.zeros(int(used.max()) + 1, dtype=np.uint8)
    for u in used.tolist():
        cen[u] = np.asarray(units[u]["centers"], dtype=float).reshape(3)
        r = units[u]["rotations"]
        if r is not None and np.asarray(r).shape[0] > 0:
            rot[u] = np.asarray(r, dtype=float).reshape(9)
            has[u] = 1
    qc, qr, qh = disc._pack_instance(querycenters, queryrotations)
    d = disc.one_vs_many(qc, qr, cen[idx], rot[idx].reshape(len(idx), n, 3, 3), cutoff=None if n == 2 else cutoff,
                         q_has_rot=qh, has_rot=has[idx], device=device)
    out = {}
    for p, v in zip(possibilities, d.tolist()):
        if v == v and v < cutoff:                      # NaN = the reference's None
            out[p] = v
    return out
