"""Ordering by similarity: the oracle restatement against the golden vectors the UNMODIFIED reference
made (tools/make_golden.py ordering), and - on a GPU - the CUDA path against both."""
import gzip
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ordering_oracle as oo  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "ordering.json.gz")


@pytest.fixture(scope="module")
def cases():
    with gzip.open(GOLD, "rt") as f:
        return json.load(f)


def test_oracle_tree_penalty_is_scipy_average_linkage(cases):
    for name, c in cases.items():
        d = np.array(c["distance"])
        assert np.array_equal(oo.treePenalty(d), np.array(c["penalty"])), name
        assert np.array_equal(oo.normalizePenaltyMatrix(d, np.array(c["penalty"])), np.array(c["normalized"])), name


def test_oracle_greedy_insertion_and_two_opt(cases):
    for name, c in cases.items():
        d = np.array(c["distance"])
        path, score = oo.greedyInsertionPathLength(d, c["start"])
        assert [path, score] == c["greedy"], name
        order, red = oo.twoOptSwap(d, path)
        assert [order, red] == c["twoopt"], name
        order, red = oo.twoOptSwap(d, c["start"])
        assert [order, red] == c["twoopt_from_start"], name


def test_oracle_tree_penalized_path_length(cases):
    for name, c in cases.items():
        d = np.array(c["distance"])
        for key, want in c["tppl"].items():
            reps, seed, strength = key.split(",")
            order = oo.treePenalizedPathLength(d, int(reps), int(seed), float(strength))
            assert order == want["order"], (name, key)
            assert oo.standardOrder(d, order) == want["standard"], (name, key)
        for key, want in c["tppl_search"].items():
            reps, seed = key.split(",")
            assert oo.treePenalizedPathLength_search(d, int(reps), int(seed)) == want, (name, key)
        assert np.array_equal(oo.reorderSymmetricMatrix(d, c["tppl"]["10,1,0.5"]["order"]), np.array(c["reordered"])), name


# ------------------------------------------------------------------------------------------------------------------
# the CUDA path (K8, csrc/ordering.cuh) through the C ABI: bit-identical to the reference's golden vectors and, on
# larger seeded matrices, to the oracle
def _gpu_modules():
    from fr3d_python_b200.ordering import orderBySimilarity as obs
    from fr3d_python_b200.search import orderBySimilarity as sobs
    return obs, sobs


@pytest.mark.gpu
def test_cuda_tree_penalty_equals_reference(cases):
    obs, _ = _gpu_modules()
    for name, c in cases.items():
        d = np.array(c["distance"])
        pen = obs.treePenalty(d)
        assert isinstance(pen, np.ndarray) and np.array_equal(pen, np.array(c["penalty"])), name
        assert np.array_equal(obs.normalizePenaltyMatrix(d, pen), np.array(c["normalized"])), name
    for bad, msg in ((np.array([[0, 1.0], [2.0, 0]]), "symmetric"), (np.array([[1.0, 1], [1, 0]]), "diagonal"),
                     (np.array([[0, np.inf], [np.inf, 0]]), "finite")):
        with pytest.raises(ValueError, match=msg):
            obs.treePenalty(np.pad(bad, ((0, 1), (0, 1))))


@pytest.mark.gpu
def test_cuda_greedy_insertion_and_two_opt_equal_reference(cases):
    obs, _ = _gpu_modules()
    for name, c in cases.items():
        d = np.array(c["distance"])
        path, score = obs.greedyInsertionPathLength(d, c["start"])
        assert [path, score] == c["greedy"], name
        order, red = obs.twoOptSwap(d, path)
        assert [order, red] == c["twoopt"], name
        order, red = obs.twoOptSwap(d, c["start"])
        assert [order, red] == c["twoopt_from_start"], name


@pytest.mark.gpu
def test_cuda_tree_penalized_path_length_equals_reference(cases):
    import random
    obs, sobs = _gpu_modules()
    for name, c in cases.items():
        d = np.array(c["distance"])
        for key, want in c["tppl"].items():
            reps, seed, strength = key.split(",")
            order = obs.treePenalizedPathLength(d, int(reps), int(seed), float(strength))
            assert order == want["order"], (name, key)
            state = random.getstate()
            assert oo.treePenalizedPathLength(d, int(reps), int(seed), float(strength)) == order
            assert random.getstate() == state                   # the global generator is consumed like the reference does
            assert obs.standardOrder(d, order) == want["standard"], (name, key)
        for key, want in c["tppl_search"].items():
            reps, seed = key.split(",")
            assert sobs.treePenalizedPathLength(d, int(reps), int(seed)) == want, (name, key)
        order = c["tppl"]["10,1,0.5"]["order"]
        assert np.array_equal(obs.reorderSymmetricMatrix(d, order), np.array(c["reordered"])), name
        z = sobs.reorderSymmetricMatrix(d, order)
        assert np.array_equal(z, np.array(c["reordered"]) * (1 - np.eye(len(d)))), name
    assert obs.treePenalizedPathLength(np.zeros((0, 0))) == [] and obs.treePenalizedPathLength(np.zeros((1, 1))) == [0]
    assert obs.treePenalizedPathLength(np.array([[0, 1.0], [1.0, 0]])) == [0, 1]


@pytest.mark.gpu
def test_cuda_ordering_of_a_device_resident_discrepancy_matrix(golden):
    """The FR3D.py:433-467 flow without the matrix leaving HBM: all-vs-all discrepancies (K6) -> tree-penalised ordering
    (both copies) -> reordered matrix; equal to the oracle run on the downloaded matrix.  n = 700 is beyond the
    reference's 300-candidate heat-map cap."""
    import torch
    from fr3d_python_b200 import synthetic
    from fr3d_python_b200.geometry import discrepancy as disc
    obs, sobs = _gpu_modules()
    av = golden("geometry.json.gz")["all_vs_all"]
    cen, rot = synthetic.make_instances(av["base_centers"], av["base_rotations"], 700, 5000)
    dm = disc.all_vs_all(cen, rot, out_device=True)
    assert isinstance(dm, torch.Tensor) and dm.is_cuda
    host = dm.cpu().numpy()
    pen = obs.treePenalty(dm)
    assert isinstance(pen, torch.Tensor) and np.array_equal(pen.cpu().numpy(), oo.treePenalty(host))
    order = sobs.treePenalizedPathLength(dm, 12, 59)
    assert order == oo.treePenalizedPathLength_search(host, 12, 59)
    order2 = obs.treePenalizedPathLength(dm, 4, 3)
    assert order2 == oo.treePenalizedPathLength(host, 4, 3)
    assert obs.standardOrder(dm, order2) == oo.standardOrder(host, order2)
    re = obs.reorderSymmetricMatrix(dm, order2)
    assert isinstance(re, torch.Tensor) and np.array_equal(re.cpu().numpy(), oo.reorderSymmetricMatrix(host, order2))
    # properties: a permutation; the ordered path is no longer than the identity path
    assert sorted(order2) == list(range(700))
    path_len = lambda o: float(sum(host[o[k], o[k + 1]] for k in range(len(o) - 1)))
    assert path_len(order2) < path_len(list(range(700)))


@pytest.mark.gpu
def test_cuda_ordering_random_matrices_against_oracle():
    """Seeded random matrices beyond the golden set (n = 3 ... 90, continuous and tie-laden integer distances, both
    copies, several repetition counts / seeds / penalty strengths): orders, scores and 2-opt reductions equal the oracle's."""
    obs, sobs = _gpu_modules()
    rng = np.random.Generator(np.random.PCG64(1650))
    for trial in range(30):
        n = int(rng.integers(3, 91))
        if trial % 3 == 0:
            pts = rng.integers(0, 5, (n, 3)).astype(float)
            d = np.abs(pts[:, None] - pts[None]).sum(-1)
        else:
            pts = rng.normal(size=(n, int(rng.integers(1, 5))))
            d = np.sqrt(((pts[:, None] - pts[None]) ** 2).sum(-1))
        d = (d + d.T) / 2
        np.fill_diagonal(d, 0.0)
        reps, seed, strength = int(rng.integers(1, 9)), int(rng.integers(1, 1000)), float(rng.choice([0.0, 0.5, 1.0, 2.5]))
        assert np.array_equal(obs.treePenalty(d), oo.treePenalty(d)), trial
        assert obs.treePenalizedPathLength(d, reps, seed, strength) == oo.treePenalizedPathLength(d, reps, seed, strength), trial
        assert sobs.treePenalizedPathLength(d, reps, seed) == oo.treePenalizedPathLength_search(d, reps, seed), trial
        start = [int(v) for v in rng.permutation(n)]
        assert list(obs.greedyInsertionPathLength(d, start)) == list(oo.greedyInsertionPathLength(d, start)), trial
        assert list(obs.twoOptSwap(d, start)) == list(oo.twoOptSwap(d, start)), trial
