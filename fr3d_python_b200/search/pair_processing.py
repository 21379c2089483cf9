"""FR3D search screening on the GPU — the entry points of fr3d/search/pair_processing.py on top of
fr3d_radius_search (include/fr3d_b200.h).

    fixed_radius_search(num_points, models, positions, max_cutoff, square_size)   pair_processing.py:106-177
    get_pairlist(Q, models, centers, alt_index=False)                             pair_processing.py:7-104

fixed_radius_search returns the reference's list: (pnt1, pnt2, distance_squared) triples in the reference's
orientation and order.  The bucket walk runs on the device (one warp per occupied bucket and neighbour); the host
only sorts the pairs into the reference's order.  get_pairlist slices that list per pair of query positions with
numpy searches instead of the reference's hand-written bisections (same indices, including their corner cases).
There is no CPU path.
"""

from collections import defaultdict

import numpy as np

from .. import _lib
from ..engine import Engine


def radius_pairs(positions, model_ids, max_cutoff, device=None):
    """(pnt1 int32[m], pnt2 int32[m], distance_squared float64[m]) in the reference's list order.
    positions [n, 3] float64 (NaN rows are skipped), model_ids int array [n]."""
    eng = Engine.get(device)
    torch = eng.torch
    positions = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
    n = len(positions)
    if n == 0:
        z = np.zeros(0, dtype=np.int32)
        return z, z.copy(), np.zeros(0)
    model_ids = np.ascontiguousarray(model_ids, dtype=np.int32)
    with torch.cuda.device(eng.device):
        dev = eng.device
        xyz = torch.from_numpy(positions).to(dev)
        mdl = torch.from_numpy(model_ids).to(dev)
        ws_bytes = int(eng.lib.fr3d_radius_search_workspace(n))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        counters = torch.zeros(8, dtype=torch.int64, device=dev)
        cap = 64 * n + 1024
        while True:
            pa = torch.empty(cap, dtype=torch.int32, device=dev)
            pb = torch.empty(cap, dtype=torch.int32, device=dev)
            pr = torch.empty(cap, dtype=torch.int32, device=dev)
            d2 = torch.empty(cap, dtype=torch.float64, device=dev)
            _lib.check(eng.lib.fr3d_radius_search(xyz.data_ptr(), mdl.data_ptr(), n, float(max_cutoff), ws.data_ptr(),
                                                  ws_bytes, pa.data_ptr(), pb.data_ptr(), pr.data_ptr(), d2.data_ptr(),
                                                  cap, counters.data_ptr(), eng.stream_ptr()), "fr3d_radius_search")
            eng.launches += 5
            cnt = counters.cpu()
            if int(cnt[2]) != 0:
                raise _lib.Fr3dError("fr3d_radius_search: coordinates outside +-2046 cutoffs (FR3D_E_RANGE)")
            m = int(cnt[0])
            if m <= cap:
                break
            cap = m + 1024                       # FR3D_E_CAPACITY: re-run with the exact count, nothing is truncated
        a, b, r, dd = (t[:m].cpu().numpy() for t in (pa, pb, pr, d2))
    order = np.lexsort((b, a, r))                # bucket of pnt1 (by its first point), stencil entry, pnt1, pnt2
    return a[order], b[order], dd[order]


def _model_ids(models):
    ids = {}
    return np.fromiter((ids.setdefault(m, len(ids)) for m in models), dtype=np.int32, count=len(models))


def fixed_radius_search(num_points, models, positions, max_cutoff, square_size=None):
    """pair_processing.py:106-177.  square_size (the span the reference uses to linearise its bucket keys) is
    accepted for signature compatibility; the device hash keys are the cell triples themselves."""
    positions = np.asarray(positions, dtype=np.float64)[:num_points]
    a, b, dd = radius_pairs(positions, _model_ids(list(models)[:num_points]), max_cutoff)
    return list(zip(a.tolist(), b.tolist(), dd.tolist()))


def get_pairlist(Q, models, centers, alt_index=False):
    """pair_processing.py:7-104."""
    pairlist = defaultdict(dict)
    npos = Q['numpositions']
    if alt_index is False:
        keys = [(i, j, i, j) for i in range(npos) for j in range(i + 1, npos)]
    else:
        keys = [(i, j, p, p + 1 + q) for p, i in enumerate(alt_index) for q, j in enumerate(alt_index[p + 1:])]
    if Q["type"] not in ("mixed", "geometric"):
        for i, j, _, _ in keys:
            pairlist[i][j] = "full"
        return pairlist
    centers = np.asarray(centers, dtype=np.float64)
    if centers.size == 0:
        return pairlist
    a, b, dd = radius_pairs(centers, _model_ids(models), Q["largestMaxRange"])
    order = np.argsort(dd, kind="stable")        # sorted(..., key=distance) is stable too (:31)
    a, b, dd = a[order], b[order], dd[order]
    both = np.empty((2 * len(a), 2), dtype=np.int64)
    both[0::2, 0] = a; both[0::2, 1] = b
    both[1::2, 0] = b; both[1::2, 1] = a
    for i, j, qi, qj in keys:
        lo, hi = Q["requireddistanceminimum"][qi][qj], Q["requireddistancemaximum"][qi][qj]
        lo2, hi2 = lo ** 2, hi ** 2
        if alt_index is False:
            # :44-75: first triple beyond the minimum ... last triple below the maximum (never before index 0)
            start = 0 if lo <= 0 else int(np.searchsorted(dd, lo2, side='right'))
            end = max(int(np.searchsorted(dd, hi2, side='left')) - 1, 0) if len(dd) else -1
            rows = both[2 * start:2 * end + 2]
        else:
            keep = np.repeat((dd > lo2) & (dd < hi2), 2)           # :78-88
            rows = np.asarray(alt_index)[both[keep]] if keep.any() else both[:0]
        pairlist[i][j] = [tuple(r) for r in rows.tolist()]
    return pairlist
