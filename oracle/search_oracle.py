"""CPU ORACLE for the FR3D search's constraint intersection — test infrastructure, NOT product code.

Restates buildSecondElementList / getPossibilities / extendFragment of fr3d/search/search.py:263-398 (the unmodified
module cannot be imported as shipped: its `from query_processing import synonym` hits an IndentationError in
query_processing.py under Python 3; tools/make_golden.py imports it with that one module stubbed).

Parity status: PINNED.  tests/test_possibilities.py compares this restatement with tests/golden/possibilities.json.gz,
made by `tools/make_golden.py possibilities` from the reference's own functions (pair lists from the unmodified
pair_processing.get_pairlist on the 1GID centres; geometric, symbolic-with-full-lists and underspecified queries).
Results are compared as sorted lists: the reference's order is the iteration order of Python sets.
"""

import numpy as np

FULL = "full"
UNDERSPECIFIED = "Query is underspecified, halting search. Add more constraints."


def build_second_elements(listOfPairs, universe):
    """search.py:377-398: sel[i][j] = {first: set(second)} or {u: "full" for u in universe[i]}."""
    k = len(listOfPairs) + 1
    sel = {}
    for i in range(k - 1):
        for j in range(i + 1, k):
            if listOfPairs[i][j] == FULL:
                sel[i, j] = {c: FULL for c in universe[i]}
            else:
                d = {}
                for a, b in listOfPairs[i][j]:
                    d.setdefault(a, set()).add(b)
                sel[i, j] = d
    return sel


def _meet(a, b):
    """search.py:111-124."""
    if a == FULL:
        return b
    if b == FULL:
        return a
    return a & b


def _distance_error(Q, units, i, j, a, b):
    """search.py:127-134."""
    return (Q["permutedDistance"][i][j] - np.linalg.norm(units[a]["centers"] - units[b]["centers"])) ** 2


def get_possibilities(Q, units, listOfPairs, universe):
    """search.py:328-375 + :263-325 -> sorted list of tuples; sets Q["halt"] / Q["errorMessage"] like the reference."""
    k = Q["numpositions"]
    if k == 2 and listOfPairs[0][1] == FULL:
        Q["errorMessage"].append(UNDERSPECIFIED)
        Q["halt"] = True
        return []
    sel = build_second_elements(listOfPairs, universe)
    geometric = Q["type"] != "symbolic"
    found = []

    def extend(fragment, ahead, err):
        # ahead[i - n] = what position i may still be, given the fragment (n = index of its last element)
        if len(fragment) == k:
            found.append(fragment)
            return
        n = len(fragment) - 1
        last = fragment[-1]
        nxt = []
        for i in range(n + 1, k):
            choices = _meet(ahead[i - n], sel[n, i][last]) if last in sel[n, i] else set()
            if len(choices) == 0:
                return
            nxt.append(choices)
        if nxt[0] == FULL:
            Q["errorMessage"].append(UNDERSPECIFIED)
            Q["halt"] = True
            return
        for c in nxt[0]:
            if not geometric:
                if c not in fragment:
                    extend(fragment + (c,), nxt, 0)
            else:
                total = err
                for j in range(0, n):                       # the last element of the fragment is not included
                    total += _distance_error(Q, units, j, n + 1, fragment[j], c)
                if total <= Q["cutoff"][n + 1]:
                    extend(fragment + (c,), nxt, total)

    for first in listOfPairs[0][1]:
        if all(first[0] in sel[0, i] for i in range(1, k)):
            ahead = [sel[0, i][first[0]] for i in range(1, k)]
            err = _distance_error(Q, units, 0, 1, first[0], first[1]) if geometric else 0
            extend(tuple(first), ahead, err)
    return sorted(found)
