"""Drop-in fr3d/ordering/orderBySimilarity.py (tree-penalised path length ordering, Aliyev & Zirbel) on the CUDA
ordering kernels (K8, csrc/ordering.cuh): same function names, arguments, return types and - bit for bit - the same
orders and scores as the reference.  `distance` may be a numpy array or a CUDA torch tensor (e.g. the matrix
geometry.discrepancy.all_vs_all(..., out_device=True) left in HBM); matrix results come back in the same form.

The starting orders are drawn on the host from Python's GLOBAL `random` generator exactly as the reference draws
them (`if seed: random.seed(seed)`, one random.shuffle per repetition), so a seeded call reproduces the reference's
ordering and leaves the generator in the same state.  There is no CPU fallback: without the CUDA library every
function that computes raises."""

import math
import random

import numpy as np

from .. import _lib
from ..engine import Engine


class _Dev:
    """A distance matrix resident on the engine's device + how to hand results back."""

    def __init__(self, distance, device=None):
        self.eng = Engine.get(device)
        torch = self.eng.torch
        self.is_tensor = isinstance(distance, torch.Tensor)
        if self.is_tensor:
            t = distance.to(device=self.eng.device, dtype=torch.float64).contiguous()
        else:
            a = np.ascontiguousarray(distance, dtype=np.float64)
            if a.ndim != 2 or a.shape[0] != a.shape[1]:
                raise ValueError("distance must be a square matrix")
            t = torch.from_numpy(a).to(self.eng.device)
        if t.dim() != 2 or t.shape[0] != t.shape[1]:
            raise ValueError("distance must be a square matrix")
        self.t = t
        self.n = int(t.shape[0])

    def back(self, tensor):
        return tensor if self.is_tensor else tensor.cpu().numpy()


def _workspace(eng, n, reps):
    nbytes = eng.lib.fr3d_order_workspace(n, reps)
    return eng.torch.empty(max(int(nbytes), 256), dtype=eng.torch.uint8, device=eng.device), int(nbytes)


def _tree_penalty(dv):
    eng, torch, n = dv.eng, dv.eng.torch, dv.n
    with torch.cuda.device(eng.device):
        penalty = torch.empty((n, n), dtype=torch.float64, device=eng.device)
        flags = torch.zeros(1, dtype=torch.int32, device=eng.device)
        ws, nbytes = _workspace(eng, n, 0)
        _lib.check(eng.lib.fr3d_tree_penalty(dv.t.data_ptr(), n, penalty.data_ptr(), flags.data_ptr(), ws.data_ptr(), nbytes,
                                             eng.stream_ptr()), "fr3d_tree_penalty")
        eng.launches += 3
        f = int(flags.item())
    # scipy.spatial.distance.squareform / scipy.cluster.hierarchy.linkage refuse these inputs the same way
    if f & 1:
        raise ValueError("Distance matrix 'X' must be symmetric.")
    if f & 2:
        raise ValueError("Distance matrix 'X' diagonal must be zero.")
    if f & 4:
        raise ValueError("The condensed distance matrix must contain only finite values.")
    return penalty


def _penalized(dv, penalty, mode, strength):
    eng, torch, n = dv.eng, dv.eng.torch, dv.n
    with torch.cuda.device(eng.device):
        out = torch.empty((n, n), dtype=torch.float64, device=eng.device)
        sums = torch.zeros(2, dtype=torch.float64, device=eng.device)
        _lib.check(eng.lib.fr3d_penalized_distance(dv.t.data_ptr(), penalty.data_ptr(), n, mode, float(strength),
                                                   sums.data_ptr(), out.data_ptr(), eng.stream_ptr()), "fr3d_penalized_distance")
        eng.launches += 2 if mode else 1
    return out


def _paths(eng, dist_t, n, starts, mode):
    """starts: list of starting orders; returns (orders [reps][n] int32 ndarray, scores, reductions as Python floats)."""
    torch = eng.torch
    reps = len(starts)
    with torch.cuda.device(eng.device):
        st = torch.from_numpy(np.ascontiguousarray(starts, dtype=np.int32).reshape(reps, n)).to(eng.device)
        orders = torch.empty((reps, n), dtype=torch.int32, device=eng.device)
        sc = torch.empty((2, reps), dtype=torch.float64, device=eng.device)
        ws, nbytes = _workspace(eng, n, reps)
        _lib.check(eng.lib.fr3d_order_paths(dist_t.data_ptr(), n, st.data_ptr(), reps, mode, orders.data_ptr(), sc[0].data_ptr(),
                                            sc[1].data_ptr(), ws.data_ptr(), nbytes, eng.stream_ptr()), "fr3d_order_paths")
        eng.launches += 1
        sc = sc.cpu().numpy()
        return orders.cpu().numpy(), sc[0].tolist(), sc[1].tolist()


def _check_order(order, n):
    order = [int(v) for v in order]
    if sorted(order) != list(range(n)):
        raise ValueError("order must be a permutation of range(%d)" % n)
    return order


def _shuffled(n, repetitions, seed):
    if seed:
        random.seed(seed)
    starts = []
    for _ in range(repetitions):
        order = list(range(0, n))
        random.shuffle(order)
        starts.append(order)
    return starts


def treePenalty(distance, device=None):
    """orderBySimilarity.py:10-41: penalty[i][j] = height of the average-linkage merge that joins i and j."""
    dv = _Dev(distance, device)
    if dv.n == 0:
        return dv.back(dv.t.clone())
    return dv.back(_tree_penalty(dv))


def normalizePenaltyMatrix(distance, penalty, device=None):
    """orderBySimilarity.py:43-64: penalty * sum(distance) / sum(penalty) (sums taken row by row, like the reference)."""
    dv = _Dev(distance, device)
    pv = _Dev(penalty, device)
    if dv.n == 0:
        return dv.back(pv.t)
    return dv.back(_penalized(dv, pv.t, 2, 1.0))


def twoOptSwap(distance, order, device=None):
    """orderBySimilarity.py:66-80 -> (order, distance_reduction)."""
    dv = _Dev(distance, device)
    order = _check_order(order, dv.n)
    if dv.n < 3:
        return order, 0
    orders, _, red = _paths(dv.eng, dv.t, dv.n, [order], 2)
    return orders[0].tolist(), red[0]


def greedyInsertionPathLength(distance, order=[], device=None):
    """orderBySimilarity.py:82-115 -> (path, score); without a starting order, a random one (global generator)."""
    dv = _Dev(distance, device)
    if len(order) == 0:
        order = list(range(0, dv.n))
        random.shuffle(order)
    order = _check_order(order, dv.n)
    orders, score, _ = _paths(dv.eng, dv.t, dv.n, [order], 1)
    return orders[0].tolist(), score[0]


def multipleGreedyInsertionPathLength(distance, repetitions=100, seed=None, device=None):
    """orderBySimilarity.py:117-129: best of `repetitions` greedy insertions (the first one on equal scores)."""
    dv = _Dev(distance, device)
    starts = _shuffled(dv.n, repetitions, seed)
    orders, scores, _ = _paths(dv.eng, dv.t, dv.n, starts, 1)
    bestScore = float("inf")
    bestPath = None
    for rep in range(0, repetitions):
        if scores[rep] < bestScore:
            bestScore = scores[rep]
            bestPath = orders[rep].tolist()
    return bestPath


def multipleGreedyInsertionPathLengthTwoOpt(distance, repetitions=10, seed=None, bestScore=float("inf"), bestOrder=[],
                                            device=None):
    """orderBySimilarity.py:131-163 -> (bestOrder, bestScore): every repetition is one CTA (greedy insertion + 2-opt),
    the best one gets a final 2-opt pass."""
    dv = _Dev(distance, device)
    n = dv.n
    if n == 0:
        return [], 0
    if n == 1:
        return [0], 0
    starts = _shuffled(n, repetitions, seed)
    if repetitions > 0:
        orders, scores, reductions = _paths(dv.eng, dv.t, n, starts, 3)
        for rep in range(0, repetitions):
            score = scores[rep] + reductions[rep]
            if score < bestScore:
                bestScore = score
                bestOrder = orders[rep].tolist()
    bestOrder, distance_reduction = twoOptSwap(dv.t, bestOrder, device)
    return bestOrder, bestScore + distance_reduction


def treePenalizedPathLength(distance, repetitions=10, seed=None, penaltyStrength=0.5, device=None):
    """orderBySimilarity.py:165-186 -> order (list of int)."""
    dv = _Dev(distance, device)
    n = dv.n
    if n == 0:
        return []
    elif n == 1:
        return [0]
    elif n == 2:
        return [0, 1]
    if penaltyStrength > 0:
        penalized = _penalized(dv, _tree_penalty(dv), 1, penaltyStrength)
        order, _ = multipleGreedyInsertionPathLengthTwoOpt(penalized, repetitions, seed, device=device)
    else:
        order, _ = multipleGreedyInsertionPathLengthTwoOpt(dv.t, repetitions, seed, device=device)
    return order


def _to_host(distance):
    return distance.cpu().numpy() if hasattr(distance, "cpu") else np.asarray(distance)


def standardOrder(distance, order):
    """orderBySimilarity.py:198-228: orient the path so that the lower discrepancies are in the upper left.  Reads
    O(n^2 / 9) entries; on the host, from the two corner blocks only."""
    n = distance.shape[0]
    if n <= 2:
        return order
    m = {3: 2, 4: 3, 5: 3, 6: 4}.get(n, int(math.ceil(n / 3)) + 1)
    head = [int(v) for v in order[:m]]
    tail = [int(order[n - i - 1]) for i in range(m)]
    if hasattr(distance, "cpu"):                       # a device tensor: only the two corner blocks cross PCIe
        idx = distance.new_tensor(head + tail).long()
        sub = distance[idx][:, idx].cpu().numpy()
        ul, lr = sub[:m, :m], sub[m:, m:]
    else:
        d = np.asarray(distance)
        ul, lr = d[np.ix_(head, head)], d[np.ix_(tail, tail)]
    upperLeftSum = 0.0
    lowerRightSum = 0.0
    for i in range(0, m):
        for j in range(i + 1, m):
            upperLeftSum += float(ul[i][j])
            lowerRightSum += float(lr[i][j])
    if upperLeftSum > lowerRightSum:
        return [order[n - i - 1] for i in range(0, n)]
    return order


def setDiagonalToZero(distance):
    """orderBySimilarity.py:230-235."""
    new_distance = distance.clone() if hasattr(distance, "clone") else np.copy(distance)
    if hasattr(new_distance, "fill_diagonal_"):
        new_distance.fill_diagonal_(0.0)
    else:
        np.fill_diagonal(new_distance, 0.0)
    return new_distance


def reorderSymmetricMatrix(distance, newOrder, device=None, zero_diagonal=False):
    """orderBySimilarity.py:237-245: out[i][j] = out[j][i] = distance[newOrder[i]][newOrder[j]] for j >= i."""
    dv = _Dev(distance, device)
    eng, torch, n = dv.eng, dv.eng.torch, dv.n
    order = _check_order(newOrder, n)
    with torch.cuda.device(eng.device):
        o = torch.tensor(order, dtype=torch.int32, device=eng.device)
        out = torch.empty((n, n), dtype=torch.float64, device=eng.device)
        _lib.check(eng.lib.fr3d_reorder_symmetric(dv.t.data_ptr(), n, o.data_ptr(), 1 if zero_diagonal else 0, out.data_ptr(),
                                                  eng.stream_ptr()), "fr3d_reorder_symmetric")
        eng.launches += 1
    return dv.back(out)


def reorderList(oldList, newOrder):
    """orderBySimilarity.py:247-253."""
    return [oldList[newOrder[i]] for i in range(0, len(oldList))]


def imputeNANValues(distance):
    """orderBySimilarity.py:255-280: NaN / negative entries become the largest finite off-diagonal value of the upper
    triangle (0 on the diagonal); the result is the symmetrised upper triangle."""
    d = _to_host(distance)
    n = d.shape[0]
    up = np.triu(d, 1)
    mask = np.triu(np.ones((n, n), dtype=bool), 1)
    vals = up[mask]
    finite = vals[~np.isnan(vals)]
    maxVal = max(0, finite.max()) if finite.size else 0
    new = np.where(np.isnan(up) | (up < 0), maxVal, up) * mask
    new = new + new.T
    diag = np.diag(d)
    np.fill_diagonal(new, np.where(np.isnan(diag) | (diag < 0), 0, diag))
    return new


def calculateDistanceMatrix(points):
    """orderBySimilarity.py:282-290 (host helper of the reference's example)."""
    n = len(points)
    distance = np.zeros((n, n))
    for i in range(0, n):
        for j in range(i + 1, n):
            distance[i][j] = np.linalg.norm(points[i] - points[j])
            distance[j][i] = distance[i][j]
    return distance
