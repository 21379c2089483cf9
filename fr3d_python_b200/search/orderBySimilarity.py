"""Drop-in fr3d/search/orderBySimilarity.py - the copy fr3d/search/FR3D.py:463 calls as
treePenalizedPathLength(allvsallmatrix, 100, 59): distance + 5 x tree penalty (:39-42), best of `repetitions` greedy
insertions (:49-103), no 2-opt - on the CUDA ordering kernels (K8), bit-identical orders.  `distance` may be a numpy
array or a CUDA tensor (the matrix geometry.discrepancy.all_vs_all left in HBM)."""

from ..ordering import orderBySimilarity as _obs
from ..ordering.orderBySimilarity import treePenalty, multipleGreedyInsertionPathLength  # noqa: F401  (same functions)


def greedyInsertionPathLength(distance, order=[], verbose=False, device=None):
    """search/orderBySimilarity.py:49-89."""
    return _obs.greedyInsertionPathLength(distance, order, device)


def treePenalizedPathLength(distance, repetitions=100, seed=None, device=None):
    """search/orderBySimilarity.py:39-42."""
    dv = _obs._Dev(distance, device)
    penalized = _obs._penalized(dv, _obs._tree_penalty(dv), 0, 5.0)
    return _obs.multipleGreedyInsertionPathLength(penalized, repetitions, seed, device)


def reorderSymmetricMatrix(distance, newOrder, device=None):
    """search/orderBySimilarity.py:105-113: upper triangle mirrored, the diagonal stays zero."""
    return _obs.reorderSymmetricMatrix(distance, newOrder, device, zero_diagonal=True)
