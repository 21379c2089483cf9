"""CPU ORACLE for the ordering-by-similarity step — test infrastructure, NOT product code.

A plain Python restatement of the tree-penalised path length ordering the FR3D search applies to the all-vs-all
discrepancy matrix: both live copies of the reference,
  fr3d/ordering/orderBySimilarity.py   (normalised penalty, greedy insertion + 2-opt; :10-186)
  fr3d/search/orderBySimilarity.py     (penalty x 5, greedy insertion only; :9-100, called at search/FR3D.py:463).
The average-linkage clustering both copies take from scipy (scipy.cluster.hierarchy.linkage(squareform(d), "average"),
scipy unpinned in setup.py) is restated here as the nearest-neighbour-chain algorithm scipy documents for the
"average" method, with the Lance-Williams update (n_x d_xi + n_y d_yi) / (n_x + n_y); only the merge HEIGHTS matter
(the penalty of a pair is the height of the merge that joins them), so the sorting / relabelling of the linkage
matrix is not needed.

Parity status: PINNED.  tests/test_ordering.py checks every function here against tests/golden/ordering.json.gz,
which `tools/make_golden.py ordering` produced by running the UNMODIFIED reference functions (and scipy's linkage) on
seeded matrices incl. tie-laden integer matrices and a discrepancy matrix of sarcin-ricin instances.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""

import math
import random

import numpy as np


def linkage_heights(distance):
    """Average linkage of a symmetric matrix with zero diagonal by nearest-neighbour chain.  Returns the list of
    merges (x, y, height, members_x, members_y) in the order the chain makes them."""
    n = distance.shape[0]
    D = np.array(distance, dtype=np.float64)
    size = [1] * n
    members = [[i] for i in range(n)]
    chain = []
    merges = []
    for _ in range(n - 1):
        if not chain:
            for i in range(n):
                if size[i] > 0:
                    chain.append(i)
                    break
        while True:
            x = chain[-1]
            if len(chain) > 1:                       # prefer the previous element of the chain on ties
                y = chain[-2]
                current_min = D[x, y]
            else:
                y = -1
                current_min = math.inf
            for i in range(n):
                if size[i] == 0 or i == x:
                    continue
                if D[x, i] < current_min:
                    current_min = D[x, i]
                    y = i
            if len(chain) > 1 and y == chain[-2]:
                break
            chain.append(y)
        chain.pop()
        chain.pop()
        if x > y:
            x, y = y, x
        nx, ny = size[x], size[y]
        merges.append((x, y, float(current_min), members[x], members[y]))
        size[x] = 0
        size[y] = nx + ny
        members[y] = members[x] + members[y]
        for i in range(n):
            if size[i] == 0 or i == y:
                continue
            v = (nx * D[i, x] + ny * D[i, y]) / (nx + ny)
            D[i, y] = v
            D[y, i] = v
    return merges


def treePenalty(distance):
    """ordering/orderBySimilarity.py:10-41 == search/orderBySimilarity.py:9-37: penalty[i][j] = height of the
    average-linkage merge that first puts i and j in one cluster."""
    distance = np.asarray(distance, dtype=np.float64)
    penalty = np.zeros(distance.shape)
    for _, _, h, a, b in linkage_heights(distance):
        for i in a:
            for j in b:
                penalty[i][j] = h
                penalty[j][i] = h
    return penalty


def normalizePenaltyMatrix(distance, penalty):
    """ordering/orderBySimilarity.py:43-64 — the two sums run row by row, left to right."""
    dsum = 0.0
    psum = 0.0
    for i in range(distance.shape[0]):
        for j in range(distance.shape[0]):
            dsum = dsum + distance[i][j]
            psum = psum + penalty[i][j]
    if psum > 0:
        return penalty * dsum / psum
    return penalty


def greedyInsertionPathLength(distance, order):
    """ordering/orderBySimilarity.py:82-115 with the starting order given (the caller shuffles)."""
    path = list(order[:2])
    score = distance[path[0], path[1]]
    for p in range(2, len(order)):
        q = order[p]
        bestScore = distance[q][path[0]]
        bestPosition = 0
        currentScore = distance[path[-1]][q]
        if currentScore < bestScore:
            bestScore = currentScore
            bestPosition = len(path)
        for position in range(1, len(path)):
            currentScore = distance[path[position - 1]][q] + distance[q][path[position]] - distance[path[position - 1]][path[position]]
            if currentScore < bestScore:
                bestScore = currentScore
                bestPosition = position
        path.insert(bestPosition, q)
        score += bestScore
    return path, score


def twoOptSwap(distance, order):
    """ordering/orderBySimilarity.py:66-80."""
    n = len(order)
    order = list(order)
    distance_reduction = 0
    for i in range(0, n):
        for p in range(n - i - 1, 1, -1):
            j = i + p
            change = distance[order[i]][order[j - 1]] + distance[order[i + 1]][order[j]] - distance[order[i]][order[i + 1]] - distance[order[j - 1]][order[j]]
            if change < -0.0000000000001:
                distance_reduction = distance_reduction + change
                order = order[:i + 1] + order[j - 1:i:-1] + order[j:]
    return order, distance_reduction


def shuffled_orders(n, repetitions, seed):
    """The starting orders the reference draws: `if seed: random.seed(seed)` once, then one random.shuffle of
    list(range(n)) per repetition, on the GLOBAL random state."""
    if seed:
        random.seed(seed)
    out = []
    for _ in range(repetitions):
        order = list(range(n))
        random.shuffle(order)
        out.append(order)
    return out


def multipleGreedyInsertionPathLengthTwoOpt(distance, repetitions=10, seed=None):
    """ordering/orderBySimilarity.py:131-163."""
    n = distance.shape[0]
    if n == 0:
        return [], 0
    if n == 1:
        return [0], 0
    bestScore = float("inf")
    bestOrder = []
    for start in shuffled_orders(n, repetitions, seed):
        order, score = greedyInsertionPathLength(distance, start)
        order, reduction = twoOptSwap(distance, order)
        score = score + reduction
        if score < bestScore:
            bestScore = score
            bestOrder = order
    bestOrder, reduction = twoOptSwap(distance, bestOrder)
    return bestOrder, bestScore + reduction


def treePenalizedPathLength(distance, repetitions=10, seed=None, penaltyStrength=0.5):
    """ordering/orderBySimilarity.py:165-186."""
    distance = np.asarray(distance, dtype=np.float64)
    n = distance.shape[0]
    if n == 0:
        return []
    if n == 1:
        return [0]
    if n == 2:
        return [0, 1]
    if penaltyStrength > 0:
        penalty = normalizePenaltyMatrix(distance, treePenalty(distance))
        order, _ = multipleGreedyInsertionPathLengthTwoOpt(distance + penaltyStrength * penalty, repetitions, seed)
    else:
        order, _ = multipleGreedyInsertionPathLengthTwoOpt(distance, repetitions, seed)
    return order


def treePenalizedPathLength_search(distance, repetitions=100, seed=None):
    """search/orderBySimilarity.py:39-42 + :91-103 (the copy search/FR3D.py:463 calls with 100, 59): distance + 5 x
    penalty, greedy insertion only, best of the repetitions (first one on equal scores)."""
    distance = np.asarray(distance, dtype=np.float64)
    penalized = distance + 5 * treePenalty(distance)
    bestScore = float("inf")
    bestPath = None
    for start in shuffled_orders(len(distance), repetitions, seed):
        path, score = greedyInsertionPathLength(penalized, start)
        if score < bestScore:
            bestScore = score
            bestPath = path
    return bestPath


def standardOrder(distance, order):
    """ordering/orderBySimilarity.py:198-228."""
    n = distance.shape[0]
    if n <= 2:
        return order
    m = {3: 2, 4: 3, 5: 3, 6: 4}.get(n, int(math.ceil(n / 3)) + 1)
    upperLeftSum = 0.0
    lowerRightSum = 0.0
    for i in range(0, m):
        for j in range(i + 1, m):
            upperLeftSum += distance[order[i]][order[j]]
            lowerRightSum += distance[order[n - i - 1]][order[n - j - 1]]
    if upperLeftSum > lowerRightSum:
        return [order[n - i - 1] for i in range(0, n)]
    return order


def reorderSymmetricMatrix(distance, newOrder):
    """ordering/orderBySimilarity.py:237-245 (diagonal included; the search copy, :105-113, leaves it zero)."""
    idx = np.asarray(newOrder, dtype=np.int64)
    g = distance[np.ix_(idx, idx)]
    return np.triu(g) + np.triu(g, 1).T
