"""FR3D search, constraint intersection (fr3d/search/search.py:263-473): the oracle restatement against the golden
vectors made by the reference's own functions (tools/make_golden.py possibilities) and - on a GPU - the CUDA path (K9)."""
import gzip
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import search_oracle as so  # noqa: E402


@pytest.fixture(scope="module")
def cases():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "possibilities.json.gz"), "rt") as f:
        return json.load(f)


def inputs(c):
    k = c["numpositions"]
    lop = {i: {} for i in range(k - 1)}
    for key, v in c["listOfPairs"].items():
        i, j = (int(x) for x in key.split(","))
        lop[i][j] = v if v == "full" else [tuple(p) for p in v]
    universe = [set(u) for u in c["universe"]]
    Q = {"type": c["type"], "numpositions": k, "MAXTIME": 1e9, "CPUTimeUsed": 0, "errorMessage": [], "halt": False}
    units = []
    if "centers" in c:
        Q["distance"] = np.array(c["distance"])
        Q["cutoff"] = c["cutoff"]
        units = [{"centers": np.array(x)} for x in c["centers"]]
    return Q, lop, universe, units


def permuted(Q, perm):
    k = Q["numpositions"]
    if "distance" in Q:
        Q["permutedDistance"] = np.zeros((k, k))
        for i in range(k):
            for j in range(k):
                Q["permutedDistance"][i][j] = Q["distance"][perm[i]][perm[j]]
    return Q


def test_host_reorder_positions_equals_reference(cases):
    from fr3d_python_b200.search import search as S          # importing the module needs no GPU
    for name, c in cases.items():
        Q, lop, universe, _ = inputs(c)
        if name == "symbolic_halt_2":
            continue
        assert S.reorderPositions(lop, universe)[2] == c["perm"], name


def test_oracle_possibilities_equal_reference(cases):
    from fr3d_python_b200.search import search as S
    for name, c in cases.items():
        Q, lop, universe, units = inputs(c)
        if name != "symbolic_halt_2":
            lop, universe, perm = S.reorderPositions(lop, universe)
            Q = permuted(Q, perm)
        got = so.get_possibilities(Q, units, lop, universe)
        assert [list(p) for p in got] == c["possibilities_sorted"], name
        assert Q["halt"] == c["halt"] and Q["errorMessage"][:1] == c["errorMessage"], name
    assert cases["symbolic_halt_3"]["halt"] and cases["symbolic_halt_2"]["halt"]


@pytest.mark.gpu
def test_cuda_possibilities_equal_reference(cases):
    from fr3d_python_b200.search import search as S
    for name, c in cases.items():
        Q, lop, universe, units = inputs(c)
        if name != "symbolic_halt_2":
            lop, universe, perm = S.reorderPositions(lop, universe)
            Q = permuted(Q, perm)
        sel = S.buildSecondElementList(Q, lop, universe, {"units": units})
        Q, got = S.getPossibilities(Q, {"units": units}, None, sel, c["numpositions"], lop)
        assert all(isinstance(p, tuple) for p in got)
        assert [list(p) for p in got] == c["possibilities_sorted"], name
        assert Q["halt"] == c["halt"] and Q["errorMessage"][:1] == c["errorMessage"], name


@pytest.mark.gpu
def test_cuda_possibilities_accept_the_reference_dict_form_and_feed_the_discrepancy_screen(cases):
    """getPossibilities called with the reference's nested-dict secondElementList, then search.py:1067-1085 on the result."""
    from fr3d_python_b200.search import search as S
    c = cases["geometric_4"]
    Q, lop, universe, units = inputs(c)
    lop, universe, perm = S.reorderPositions(lop, universe)
    Q = permuted(Q, perm)
    sel = {i: {j: v for (a, j), v in so.build_second_elements(lop, universe).items() if a == i} for i in range(3)}
    Q, got = S.getPossibilities(Q, {"units": units}, perm, sel, 4, lop)
    assert [list(p) for p in got] == c["possibilities_sorted"]
    assert tuple(c["motif"][p] for p in perm) in set(got)          # the query motif itself is among its possibilities


@pytest.mark.gpu
def test_cuda_possibilities_random_queries_against_oracle():
    """Seeded random queries beyond the golden set: symbolic queries of 3-6 positions with random pair lists, random "full"
    lists and partial universes; geometric queries on random point clouds with tight and loose cutoffs (so that the
    running distance-error test, incl. its skipped last term, decides many fragments)."""
    from fr3d_python_b200.search import search as S
    rng = np.random.Generator(np.random.PCG64(2630))
    n_checked = 0
    for trial in range(24):
        k = int(rng.integers(3, 7))
        U = int(rng.integers(30, 90))
        symbolic = trial % 2 == 0
        lop = {i: {} for i in range(k - 1)}
        universe = [set(int(v) for v in rng.choice(U, int(0.8 * U), replace=False)) for _ in range(k)]
        units, Q = [], {"type": "symbolic" if symbolic else "geometric", "numpositions": k, "errorMessage": [], "halt": False}
        if symbolic:
            for i in range(k - 1):
                for j in range(i + 1, k):
                    if rng.random() < 0.25 and not (i == 0 and j == 1):
                        lop[i][j] = "full"
                    else:
                        m = int(rng.integers(U, 6 * U))
                        a, b = rng.integers(0, U, m), rng.integers(0, U, m)
                        lop[i][j] = sorted(set((int(x), int(y)) for x, y in zip(a, b) if x != y))
        else:
            pts = rng.uniform(0, 30, (U, 3))
            units = [{"centers": p} for p in pts]
            motif = [int(v) for v in rng.choice(U, k, replace=False)]
            dist = np.linalg.norm(pts[motif][:, None] - pts[motif][None], axis=2)
            width = float(rng.uniform(4, 9))
            for i in range(k - 1):
                for j in range(i + 1, k):
                    d = np.linalg.norm(pts[:, None] - pts[None], axis=2)
                    ok = (np.abs(d - dist[i, j]) < width) & ~np.eye(U, dtype=bool)
                    lop[i][j] = [(int(a), int(b)) for a, b in zip(*np.nonzero(ok))]
            Q["distance"] = dist
            Q["cutoff"] = [float(rng.uniform(5, 40)) * (m + 1) for m in range(k)]
        lop, universe, perm = S.reorderPositions(lop, universe)
        Q = permuted(Q, perm)
        Qo = dict(Q, errorMessage=[], halt=False)
        want = so.get_possibilities(Qo, units, lop, universe)
        if Qo["halt"]:
            continue
        sel = S.buildSecondElementList(Q, lop, universe, {"units": units})
        Q, got = S.getPossibilities(Q, {"units": units}, perm, sel, k, lop)
        assert got == want, (trial, k, symbolic, len(got), len(want))
        assert not Q["halt"]
        n_checked += len(want) > 0
    assert n_checked >= 12
