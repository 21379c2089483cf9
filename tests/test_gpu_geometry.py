"""Kabsch (K5) and discrepancy (K6) kernels against the reference's known answers, the golden
outputs of the unmodified reference, and size-independent properties at larger N."""

import numpy as np
import pytest

from fr3d_python_b200 import _lib, synthetic
from fr3d_python_b200.geometry import angleofrotation, discrepancy as disc, superpositions as sup
from oracle import fr3d_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-9
NOISE_FLOOR = 1e-7     # discrepancies below this are arccos round-off in the reference itself


def test_besttransformation_known_answers(golden):
    """tests/superpositions_test.py:11-119 + reference outputs on seeded problems."""
    geo = golden("geometry.json.gz")
    by = {k["name"]: k for k in geo["kabsch"]}
    r = sup.besttransformation(by["octahedron_90z"]["set1"], by["octahedron_90z"]["set2"])
    assert len(r) == 6 and isinstance(r[0], np.matrix) and isinstance(r[1], np.matrix)
    np.testing.assert_almost_equal(np.asarray(r[0]), [[0, 1, 0], [-1, 0, 0], [0, 0, 1]], decimal=7)
    for ax in "xyz":
        U = sup.besttransformation(by["p4_45" + ax]["set1"], by["p4_45" + ax]["set2"])[0]
        np.testing.assert_almost_equal(angleofrotation.angle_of_rotation(U), np.pi / 4, decimal=7)
    for k in geo["kabsch"]:
        U, new1, mean1, rmsd, sse, mean2 = sup.besttransformation(np.array(k["set1"]), np.array(k["set2"]))
        assert np.abs(np.asarray(U) - np.array(k["U"])).max() < TOL, k["name"]
        assert np.abs(np.asarray(new1) - np.array(k["new1"])).max() < TOL      # "fitted coordinates" = dev1 . U
        assert np.abs(mean1 - np.array(k["mean1"])).max() < TOL and np.abs(mean2 - np.array(k["mean2"])).max() < TOL
        assert abs(sse - k["sse"]) < TOL and abs(rmsd - k["rmsd"]) < TOL
        r = sup.besttransformation_weighted(np.array(k["set1"]), np.array(k["set2"]), k["weights"])
        assert len(r) == 5
        assert np.abs(np.asarray(r[0]) - np.array(k["Uw"])).max() < TOL and abs(r[4] - k["ssew"]) < TOL
        # weights of the wrong length are ignored (superpositions.py:101-104)
        r1 = sup.besttransformation_weighted(np.array(k["set1"]), np.array(k["set2"]))
        assert np.abs(np.asarray(r1[0]) - np.array(k["U"])).max() < TOL
    with pytest.raises(AssertionError):
        sup.besttransformation([[0, 0, 0.0]], [[0, 0, 0.0], [1, 1, 1.0]])
    with pytest.raises(AssertionError):
        sup.besttransformation([], [])


def test_batched_kabsch_matches_oracle():
    rng = np.random.Generator(np.random.PCG64(11))
    s1, s2, ws = [], [], []
    for k in range(500):
        n = int(rng.integers(3, 40))
        a = rng.normal(size=(n, 3)) * 8
        Q = synthetic.random_rotation(rng)
        s1.append(a)
        s2.append(a @ Q.T + rng.normal(size=(n, 3)) * 0.4 + rng.normal(size=3) * 30)
        ws.append(rng.uniform(0.2, 2.0, size=n))
    r = sup.besttransformation_batched(s1, s2)
    rw = sup.besttransformation_batched(s1, s2, ws, want_new1=False)
    for k in range(500):
        U, new1, mean1, rmsd, sse, mean2 = O.besttransformation(s1[k], s2[k])
        assert np.abs(r["U"][k] - U).max() < TOL and np.abs(r["new1"][k] - new1).max() < TOL
        assert abs(r["sse"][k] - sse) < TOL * max(1.0, sse) and abs(r["rmsd"][k] - rmsd) < TOL
        Uw = O.besttransformation_weighted(s1[k], s2[k], ws[k])[0]
        assert np.abs(rw["U"][k] - Uw).max() < TOL
    assert np.abs(np.linalg.det(r["U"]) - 1).max() < 1e-12


def test_batched_kabsch_short_and_mixed_point_sets():
    """The staged form of K5 (a warp's 32 problems copied into shared memory when they have at most 352 rows in total): short
    point sets (3-12 points: staged), mixed with a few long ones (their warps work from global memory), a batch that is
    not a multiple of 32, weights - everything against the oracle, and against the same problems run one by one."""
    rng = np.random.Generator(np.random.PCG64(12))
    s1, s2, ws = [], [], []
    for k in range(1013):
        n = int(rng.integers(3, 13)) if k % 97 else int(rng.integers(60, 200))
        a = rng.normal(size=(n, 3)) * 6
        Q = synthetic.random_rotation(rng)
        s1.append(a)
        s2.append(a @ Q.T + rng.normal(size=(n, 3)) * 0.3 + rng.normal(size=3) * 20)
        ws.append(rng.uniform(0.2, 2.0, size=n))
    r = sup.besttransformation_batched(s1, s2)
    rw = sup.besttransformation_batched(s1, s2, ws)
    for k in range(0, 1013, 7):
        U, new1, mean1, rmsd, sse, mean2 = O.besttransformation(s1[k], s2[k])
        assert np.abs(r["U"][k] - U).max() < TOL and np.abs(r["new1"][k] - new1).max() < TOL
        assert np.abs(r["mean1"][k] - mean1).max() < TOL and np.abs(r["mean2"][k] - mean2).max() < TOL
        assert abs(r["sse"][k] - sse) < TOL * max(1.0, sse) and abs(r["rmsd"][k] - rmsd) < TOL
        Uw, new1w = O.besttransformation_weighted(s1[k], s2[k], ws[k])[:2]
        assert np.abs(rw["U"][k] - Uw).max() < TOL and np.abs(rw["new1"][k] - new1w).max() < TOL
    one = sup.besttransformation_batched(s1[500:501], s2[500:501])
    assert np.array_equal(one["U"][0], r["U"][500]) and np.array_equal(one["new1"][0], r["new1"][500])


def _args(k):
    return ([np.array(c) for c in k["centers1"]], [np.array(r) for r in k["rotations1"]],
            [np.array(c) for c in k["centers2"]], [np.array(r) for r in k["rotations2"]])


def test_matrix_discrepancy_known_answers(golden):
    """tests/discrepancy_tests.py:301-348."""
    geo = golden("geometry.json.gz")
    by = {k["name"]: k for k in geo["discrepancy_kats"]}
    np.testing.assert_almost_equal(0.9402, disc.matrix_discrepancy(*_args(by["simple"])), decimal=3)
    np.testing.assert_almost_equal(1.1414, disc.matrix_discrepancy(*_args(by["another"])), decimal=3)
    for k in geo["discrepancy_kats"]:
        assert abs(disc.matrix_discrepancy(*_args(k)) - k["d"]) < TOL, k["name"]
    k = by["simple"]
    assert abs(disc.matrix_discrepancy_cutoff(*_args(k), 2.5) - k["cutoff_2.5"]) < TOL
    assert disc.matrix_discrepancy_cutoff(*_args(k), 0.5) is None
    k = by["pair2"]                                    # n == 2 closed form ignores the cutoff
    assert abs(disc.matrix_discrepancy_cutoff(*_args(k), 0.5) - k["cutoff_0.5"]) < TOL
    with pytest.raises(AssertionError):
        disc.matrix_discrepancy([], [], [], [])
    a = list(_args(by["another"]))
    a[2] = a[2][:-1]
    a[3] = a[3][:-1]
    with pytest.raises(AssertionError):
        disc.matrix_discrepancy(*a)
    # np.matrix rotations (what Component.rotation_matrix is) and empty / None rotations
    a = list(_args(by["simple"]))
    a[1] = [np.matrix(r) for r in a[1]]
    assert abs(disc.matrix_discrepancy(*a) - by["simple"]["d"]) < TOL
    a = list(_args(by["simple"]))
    a[1] = [np.empty((0, 0))] + a[1][1:]
    a[3] = [None] + a[3][1:]
    want = O.matrix_discrepancy(a[0], a[1], a[2], [np.empty((0, 0))] + a[3][1:])
    assert abs(disc.matrix_discrepancy(*a) - want) < TOL
    # weights
    w = [1.0, 2.0, 0.5, 1.5, 1.0]
    a = _args(by["simple"])
    assert abs(disc.matrix_discrepancy(*a, angle_weight=2.5, center_weight=w)
               - O.matrix_discrepancy(*a, angle_weight=2.5, center_weight=w)) < TOL


def test_all_vs_all_matches_reference_block(golden):
    av = golden("geometry.json.gz")["all_vs_all"]
    C, R = synthetic.make_instances(av["base_centers"], av["base_rotations"], av["N"], av["seed"])
    assert np.array_equal(C, np.array(av["centers"])) and np.array_equal(R, np.array(av["rotations"]))
    D = disc.all_vs_all(C, R)
    assert np.abs(D - np.array(av["D"])).max() < TOL
    assert np.array_equal(D, D.T) and np.all(np.diag(D) == 0)
    has = np.ones((av["N"], av["n"]), dtype=np.uint8)
    has[:, 0] = 0
    assert np.abs(disc.all_vs_all(C, R, has_rot=has) - np.array(av["D_norot0"])).max() < TOL
    # row blocks (multi-GPU sharding) give the same numbers
    blk = disc.all_vs_all(C, R, rows=(10, 37))
    assert np.array_equal(blk, D[10:37])
    # one-vs-many with the early-exit cutoff (search.py:1073-1085)
    row = disc.one_vs_many(C[0], R[0], C, R, cutoff=av["cutoff"])
    ref = np.array([np.nan if v is None else v for v in av["row0_cutoff"]])
    assert np.array_equal(np.isnan(row), np.isnan(ref))
    ok = ~np.isnan(ref) & (ref > NOISE_FLOOR)
    assert np.abs(row[ok] - ref[ok]).max() < TOL
    listed = [(j, row[j], ref[j]) for j in np.nonzero(~np.isnan(ref) & (ref <= NOISE_FLOOR))[0]]
    assert all(abs(a) < NOISE_FLOOR for _, a, _ in listed)          # the exempted self pair


@pytest.mark.parametrize("n", [2, 3, 5, 7, 16])
def test_all_vs_all_other_sizes_against_oracle(n):
    rng = np.random.Generator(np.random.PCG64(100 + n))
    N = 24
    C = rng.normal(size=(N, n, 3)) * 6
    R = np.array([[synthetic.random_rotation(rng) for _ in range(n)] for _ in range(N)])
    D = disc.all_vs_all(C, R)
    want = O.all_vs_all_discrepancy(C, R)
    assert np.abs(D - want).max() < TOL


def test_all_vs_all_properties_at_scale():
    """N = 3000, n = 7 (9 M pairs): symmetry, zero diagonal, invariance under a rigid motion of
    every instance, self-discrepancy at the noise floor, spot check against the oracle."""
    k = synthetic.load_golden_spec  # noqa: F841  (fixtures come from the golden geometry file)
    import gzip
    import json
    import os
    from conftest import GOLD
    with gzip.open(os.path.join(GOLD, "geometry.json.gz"), "rt") as f:
        av = json.load(f)["all_vs_all"]
    N = 3000
    C, R = synthetic.make_instances(av["base_centers"], av["base_rotations"], N, 5000)
    D = disc.all_vs_all(C, R)
    assert D.shape == (N, N) and np.array_equal(D, D.T) and np.all(np.diag(D) == 0) and np.all(np.isfinite(D))
    rng = np.random.Generator(np.random.PCG64(9))
    for (i, j) in rng.integers(0, N, size=(40, 2)):
        if i != j:
            want = O.matrix_discrepancy(list(C[i]), list(R[i]), list(C[j]), list(R[j]))
            assert abs(D[i, j] - want) < TOL
    Q = synthetic.random_rotation(rng)
    C2 = C @ Q.T + np.array([3.0, -7.0, 11.0])
    R2 = np.einsum('ij,nkjl->nkil', Q, R)
    D2 = disc.all_vs_all(C2, R2)
    big = D > NOISE_FLOOR
    assert np.abs(D2 - D)[big].max() < 1e-8
    # an instance against a rigidly moved copy of itself: discrepancy at the arccos noise floor
    Dself = disc.one_vs_many(C[0], R[0], C2[:50], R2[:50])
    assert Dself[0] < NOISE_FLOOR and np.all(Dself[1:] > 0.01)


def test_argument_checks():
    C = np.zeros((4, 17, 3))
    R = np.tile(np.eye(3), (4, 17, 1, 1))
    with pytest.raises(_lib.Fr3dError):
        disc.all_vs_all(C, R)                     # more than 16 positions


def test_upper_triangle_row_blocks_and_assembly_equal_the_full_launch(golden):
    """The multi-GPU split of the matrix (parallel.triangle_blocks + fr3d_discrepancy_rows_upper +
    fr3d_discrepancy_assemble, here with 3 emulated ranks on one GPU) gives the bits of the full launch."""
    import torch
    from fr3d_python_b200 import _lib, parallel, synthetic
    from fr3d_python_b200.engine import Engine
    from fr3d_python_b200.geometry import discrepancy as disc
    av = golden("geometry.json.gz")["all_vs_all"]
    N = 1000
    C, R = synthetic.make_instances(av["base_centers"], av["base_rotations"], N, 77)
    n = C.shape[1]
    full = disc.all_vs_all(C, R)
    eng = Engine.get()
    dc = torch.from_numpy(C).to(eng.device)
    dr = torch.from_numpy(R.reshape(N, n, 9)).to(eng.device)
    D = torch.zeros((N, N), dtype=torch.float64, device=eng.device)
    parts = parallel.triangle_blocks(N, 3, 128)
    assert sorted(b for p in parts for b in p) == [(r, min(N, r + 128)) for r in range(0, N, 128)]
    for part in parts:
        for r0, r1 in part:
            blk = torch.full(((r1 - r0) * (N - r0),), -1.0, dtype=torch.float64, device=eng.device)
            _lib.check(eng.lib.fr3d_discrepancy_rows_upper(dc.data_ptr(), dr.data_ptr(), None, N, n, 1.0, r0, r1,
                                                           blk.data_ptr(), eng.stream_ptr()), "rows_upper")
            assert float(blk.min()) >= 0.0                              # every entry of the block was written
            _lib.check(eng.lib.fr3d_discrepancy_assemble(blk.data_ptr(), N, r0, r1, D.data_ptr(), eng.stream_ptr()),
                       "assemble")
    got = D.cpu().numpy()
    assert np.array_equal(got, full) and np.array_equal(got, got.T) and np.all(np.diag(got) == 0)
    # the class bench.py uses, single rank
    sh = disc.AllVsAllSharded(C, R)
    sh.upload()
    assert np.array_equal(sh.run().cpu().numpy(), full)
    assert np.array_equal(disc.all_vs_all_multi_gpu(C, R), full)
    assert disc.fp64_peak_tflops() > 5.0


def test_possibility_screen_of_the_fr3d_search(golden):
    """search.possibility_discrepancies (one launch) against the loop of fr3d/search/search.py:1067-1085 restated with
    the oracle's matrix_discrepancy_cutoff: same keys in the same order, values within 1e-9; one unit without a
    rotation matrix (NA_protein_annotation.py:1705-1714 style)."""
    from fr3d_python_b200 import synthetic
    from fr3d_python_b200.search import search as S
    from oracle import fr3d_oracle as O
    av = golden("geometry.json.gz")["all_vs_all"]
    C, R = synthetic.make_instances(av["base_centers"], av["base_rotations"], 60, 11)
    n = C.shape[1]
    units = []
    for k in range(60):
        for p in range(n):
            units.append({"centers": C[k, p], "rotations": R[k, p]})
    units[5]["rotations"] = np.array([])
    rng = np.random.default_rng(2)
    possibilities = [tuple(range(k * n, (k + 1) * n)) for k in range(60)]
    possibilities += [tuple(rng.choice(len(units), n, replace=False).tolist()) for _ in range(200)]
    qc = [C[0, p] + 0.05 for p in range(n)]
    qr = [R[0, p] for p in range(n)]
    cutoff = 1.2
    want = {}
    for poss in possibilities:
        d = O.matrix_discrepancy_cutoff(qc, qr, [units[i]["centers"] for i in poss], [units[i]["rotations"] for i in poss], cutoff)
        if d is not None and d < cutoff:
            want[poss] = d
    got = S.possibility_discrepancies(qc, qr, units, possibilities, cutoff)
    assert list(got.keys()) == list(want.keys()) and 5 < len(want) < len(possibilities)
    exempt = []
    for k in want:
        if want[k] <= NOISE_FLOOR:            # the query against its own rigid copy: arccos round-off in the reference itself
            assert got[k] < NOISE_FLOOR
            exempt.append((k, got[k], want[k]))
        else:
            assert abs(got[k] - want[k]) < 1e-9
    assert len(exempt) <= 1
    assert S.possibility_discrepancies(qc, qr, units, [], cutoff) == {}
