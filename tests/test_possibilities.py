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
