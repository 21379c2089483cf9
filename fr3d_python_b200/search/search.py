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


# ---------------------------------------------------------------------------------------------------------------------
# Constraint intersection (search.py:263-398) on the GPU: K9, csrc/possibilities.cuh
import ctypes as _C

from .. import _lib
from ..engine import Engine

MAX_POSITIONS = _lib.SEARCH_MAX_POSITIONS


class SecondElementList:
    """buildSecondElementList (search.py:377-398) in device form: for every position pair a < b the pairs of
    listOfPairs[a][b] as CSR rows "first element -> ascending second elements" (the reference's dict of sets), or the
    membership of universe[a] where the list is "full"."""

    def __init__(self, listOfPairs, universe, n_units=None, device=None):
        self.eng = Engine.get(device)
        torch = self.eng.torch
        k = len(listOfPairs) + 1
        if k > MAX_POSITIONS:
            raise _lib.Fr3dError("at most %d positions are supported" % MAX_POSITIONS)
        self.numpositions = k
        host = {}
        top = -1
        for a in range(k - 1):
            for b in range(a + 1, k):
                L = listOfPairs[a][b]
                if isinstance(L, str):                          # "full"
                    host[(a, b)] = None
                    continue
                arr = np.asarray(list(L), dtype=np.int64).reshape(-1, 2)
                host[(a, b)] = arr
                if len(arr):
                    top = max(top, int(arr.max()))
        for u in universe:
            if len(u):
                top = max(top, int(max(u)))
        self.n_units = (top + 1) if n_units is None else int(n_units)
        U = self.n_units
        self.keep = []
        self.first_pairs = None
        self.rowptr, self.cols, self.universe = {}, {}, [None] * k
        with torch.cuda.device(self.eng.device):
            for (a, b), arr in host.items():
                if arr is None:
                    continue
                key = np.unique(arr[:, 0] * (U + 1) + arr[:, 1]) if len(arr) else np.zeros(0, dtype=np.int64)
                first, second = key // (U + 1), key % (U + 1)
                rp = np.zeros(U + 1, dtype=np.int32)
                np.cumsum(np.bincount(first, minlength=U), out=rp[1:])
                self.rowptr[(a, b)] = torch.from_numpy(rp).to(self.eng.device)
                self.cols[(a, b)] = torch.from_numpy(second.astype(np.int32)).to(self.eng.device)
                if (a, b) == (0, 1):
                    self.first_pairs = torch.from_numpy(np.stack([first, second], 1).astype(np.int32)).to(self.eng.device)
            for a in range(k):
                if any(host.get((a, b), 0) is None for b in range(a + 1, k)):
                    m = np.zeros(max(U, 1), dtype=np.uint8)
                    m[np.asarray(list(universe[a]), dtype=np.int64)] = 1
                    self.universe[a] = torch.from_numpy(m).to(self.eng.device)
        self.first_full = host.get((0, 1), 0) is None

    def query(self, Q, units):
        """The device copy of fr3d_search_query_t for this list set and query."""
        torch, k, K = self.eng.torch, self.numpositions, MAX_POSITIONS
        q = _lib.SearchQueryStruct()
        q.n_positions, q.n_units = k, self.n_units
        q.symbolic = 1 if Q["type"] == "symbolic" else 0
        for (a, b), t in self.rowptr.items():
            q.rowptr[a * K + b] = t.data_ptr()
            q.cols[a * K + b] = self.cols[(a, b)].data_ptr()
        for a, t in enumerate(self.universe):
            if t is not None:
                q.universe[a] = t.data_ptr()
        centers = None
        if not q.symbolic:
            cen = np.zeros((max(self.n_units, 1), 3))
            have = units.keys() if isinstance(units, dict) else range(len(units))
            for u in have:
                if 0 <= u < self.n_units:
                    c = np.asarray(units[u]["centers"], dtype=float).reshape(-1)
                    if c.shape[0] == 3:
                        cen[u] = c
            centers = torch.from_numpy(cen).to(self.eng.device)
            q.centers = centers.data_ptr()
            pd = np.asarray(Q["permutedDistance"], dtype=float)
            for a in range(k):
                for b in range(k):
                    q.perm_dist[a * K + b] = float(pd[a][b])
            for m in range(k):
                q.cutoff[m] = float(Q["cutoff"][m])
        raw = np.frombuffer(bytes(q), dtype=np.uint8).copy()
        return torch.from_numpy(raw).to(self.eng.device), centers


def buildSecondElementList(Q, listOfPairs, universe, ifedata=None, device=None):
    """search.py:377-398."""
    return SecondElementList(listOfPairs, universe, device=device)


UNDERSPECIFIED = "Query is underspecified, halting search. Add more constraints."


def getPossibilities(Q, ifedata, perm, secondElementList, numpositions, listOfPairs, device=None):
    """search.py:328-375 -> (Q, possibilities): every tuple of unit indices that satisfies all pairwise constraints (and,
    for geometric / mixed queries, the running distance-error cutoffs).  Same SET as the reference; the order is
    lexicographic (the reference's is the iteration order of Python sets).  Q["MAXTIME"] does not truncate anything here."""
    if not isinstance(secondElementList, SecondElementList):      # the reference's nested dicts: only the universes are read
        universe = []                                             # keys of a "full" entry = universe[a] (search.py:388-391)
        for a in range(numpositions):
            u = set()
            for b in range(a + 1, numpositions):
                if isinstance(listOfPairs[a][b], str):
                    u = set(secondElementList[a][b].keys())
            universe.append(u)
        secondElementList = SecondElementList(listOfPairs, universe, device=device)
    sel = secondElementList
    if Q["numpositions"] == 2 and sel.first_full:
        Q["errorMessage"].append(UNDERSPECIFIED)
        Q["halt"] = True
        return Q, []
    if sel.first_full:
        raise _lib.Fr3dError("listOfPairs[0][1] is 'full': reorderPositions puts the shortest list first")
    eng, torch, lib = sel.eng, sel.eng.torch, sel.eng.lib
    symbolic = Q["type"] == "symbolic"
    with torch.cuda.device(eng.device):
        dq, centers = sel.query(Q, ifedata["units"] if ifedata is not None else None)
        pairs = sel.first_pairs
        n = int(pairs.shape[0])
        flags = torch.zeros(1, dtype=torch.int32, device=eng.device)
        counts = torch.empty(n + 1, dtype=torch.int64, device=eng.device)
        _lib.check(lib.fr3d_possibilities_first(dq.data_ptr(), pairs.data_ptr(), n, 0, counts.data_ptr(), None, None, 0,
                                                eng.stream_ptr()), "fr3d_possibilities_first")
        total = int(counts[n].item())
        frags = torch.empty((max(total, 1), 2), dtype=torch.int32, device=eng.device)
        errs = None if symbolic else torch.empty(max(total, 1), dtype=torch.float64, device=eng.device)
        _lib.check(lib.fr3d_possibilities_first(dq.data_ptr(), pairs.data_ptr(), n, 1, counts.data_ptr(), frags.data_ptr(),
                                                errs.data_ptr() if errs is not None else None, total, eng.stream_ptr()),
                   "fr3d_possibilities_first")
        eng.launches += 3
        for level in range(1, numpositions - 1):
            if total == 0:
                break
            counts = torch.empty(total + 1, dtype=torch.int64, device=eng.device)
            args = (dq.data_ptr(), level, frags.data_ptr(), errs.data_ptr() if errs is not None else None, total)
            _lib.check(lib.fr3d_possibilities_extend(*args, 0, counts.data_ptr(), None, None, 0, flags.data_ptr(),
                                                     eng.stream_ptr()), "fr3d_possibilities_extend")
            new_total = int(counts[total].item())
            new_frags = torch.empty((max(new_total, 1), level + 2), dtype=torch.int32, device=eng.device)
            new_errs = None if symbolic else torch.empty(max(new_total, 1), dtype=torch.float64, device=eng.device)
            _lib.check(lib.fr3d_possibilities_extend(*args, 1, counts.data_ptr(), new_frags.data_ptr(),
                                                     new_errs.data_ptr() if new_errs is not None else None, new_total,
                                                     flags.data_ptr(), eng.stream_ptr()), "fr3d_possibilities_extend")
            eng.launches += 3
            frags, errs, total = new_frags, new_errs, new_total
        if int(flags.item()) & 1:
            Q["errorMessage"].append(UNDERSPECIFIED)
            Q["halt"] = True
            return Q, []
        out = frags[:total].cpu().numpy()
    return Q, [tuple(int(v) for v in row) for row in out]


def permuteListOfPairs(listOfPairs, perm):
    """search.py:401-428: the lists re-keyed by the new position numbers, pairs flipped where the order of the two
    positions changes."""
    where = {p: i for i, p in enumerate(perm)}
    new = {}
    for i in listOfPairs:
        for j in listOfPairs[i]:
            L = listOfPairs[i][j]
            a, b = where[i], where[j]
            if a > b:
                a, b = b, a
                if not isinstance(L, str):
                    L = [(p[1], p[0]) for p in L]
            elif not isinstance(L, str):
                L = list(L)
            new.setdefault(a, {})[b] = L
    return new


def reorderPositions(listOfPairs, universe):
    """search.py:431-473: start from the shortest list, then always add the position whose lists to the chosen ones are
    shortest in total ("full" counts 10^9) -> (newListOfPairs, newUniverse, perm)."""
    k = len(listOfPairs) + 1
    lengths = np.zeros((k, k))
    sizes = []
    for i in listOfPairs:
        for j in listOfPairs[i]:
            L = 1000000000 if isinstance(listOfPairs[i][j], str) else len(listOfPairs[i][j])
            lengths[i][j] = lengths[j][i] = L
            sizes.append(([i, j], L))
    perm = sorted(sizes, key=lambda x: x[1])[0][0]
    objects = set(range(k)) - set(perm)
    while len(perm) < k:
        best_len, best = float("inf"), -1
        for obj in objects:
            total = 1
            for p in perm:
                total += lengths[p][obj]
            if total < best_len:
                best_len, best = total, obj
        perm.append(best)
        objects.remove(best)
    return permuteListOfPairs(listOfPairs, perm), [universe[i] for i in perm], perm
