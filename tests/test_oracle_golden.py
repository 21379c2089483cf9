"""The oracle (oracle/fr3d_oracle.py) against the golden vectors produced by the unmodified
reference (tools/make_golden.py) and against the known answers the reference's own tests hold.
CPU only."""

import numpy as np
import pytest

from oracle import fr3d_oracle as O


def _close_rows(a, b, tol=0.0):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x[:3] == y[:3], (x, y)
        for u, v in zip(x[3:], y[3:]):
            assert (u is None) == (v is None), (x, y)
            if u is not None:
                if isinstance(u, str):
                    assert u == v
                else:
                    assert abs(u - v) <= tol, (x, y)


@pytest.mark.parametrize("spec_name,annot_name", [
    ("spec_1s72_strand.json.gz", "annot_1s72_strand_all9.json.gz"),
    ("spec_1gid_full.json.gz", "annot_1gid_full_all9.json.gz"),
    ("spec_1gid_base.json.gz", "annot_1gid_base_bp_stack_cp.json.gz"),
    ("spec_1gid_base.json.gz", "annot_1gid_base_bp.json.gz"),
    # DNA, missing atoms, alt-id twins, insertion codes, a second model and a symmetry copy (tools/make_golden.py variants)
    ("spec_1gid_variants.json.gz", "annot_1gid_variants_all9.json.gz"),
])
def test_annotation_matches_reference(golden, spec_name, annot_name):
    spec = golden(spec_name)
    g = golden(annot_name)
    out, c2i, tr = O.annotate(spec, g["categories"])
    if g["traces"]["candidates"]:
        assert tr["candidates"] == g["traces"]["candidates"]          # oriented pairs, reference visit order
    for key in ("sO", "stacking", "sugar_ribose", "backbone"):
        assert tr[key] == g["traces"][key], key
    _close_rows(tr["coplanar"], g["traces"]["coplanar"])
    _close_rows(tr["basepair"], g["traces"]["basepair"])
    got = {k: [list(t) for t in v] for k, v in out.items()}
    assert got == g["interaction_to_list_of_tuples"]                   # same rows, same order
    assert {k: sorted(v) for k, v in c2i.items()} == g["category_to_interactions"]
    assert "Found %d nucleotide" % tr["count_pair"] in g["found_line"]


def test_1gid_probe_counts(golden):
    """basepair-only run on the 1GID bases."""
    g = golden("annot_1gid_base_bp.json.gz")
    counts = {k: len(v) for k, v in g["interaction_to_list_of_tuples"].items()}
    assert counts["cWW"] == 202 and counts["tHS"] == 7 and counts["ncWW"] == 10 and counts["p_1"] == 314
    assert "Found 154" in g["found_line"]


@pytest.mark.parametrize("name", ["1gid_full", "1s72_strand", "1s72_sarcin", "1gid_variants"])
def test_frames_match_reference(golden, name):
    spec = golden("spec_%s.json.gz" % name)
    frames = golden("frames_%s.json.gz" % name)
    nts = O.build_nts(spec)
    for nt, f in zip(nts, frames):
        assert nt.unit_id() == f["unit_id"]
        assert np.abs(nt.base_center - np.array(f["center"])).max() == 0.0
        assert np.abs(nt.rotation - np.array(f["rotation"])).max() == 0.0
        placed = {n: x for n, x in zip(nt.atom_names, nt.atom_xyz) if n in f["hydrogens"]}
        assert set(placed) == set(f["hydrogens"])
        for n, x in placed.items():
            assert np.abs(x - np.array(f["hydrogens"][n])).max() == 0.0


def test_frames_match_matlab_golden(golden):
    """tests/FR3D-Matlab-1GID-rotation-matrices.txt (second witness; inputs have 6 decimals)."""
    spec = golden("spec_1gid_base.json.gz")
    mat = golden("matlab_1gid_frames.json.gz")
    nts = O.build_nts(spec)
    assert len(nts) == len(mat) == 316
    for nt, m in zip(nts, mat):
        assert (nt.chain, nt.sequence, nt.number) == (m["chain"], m["sequence"], m["number"])
        assert np.abs(nt.base_center - np.array(m["center"])).max() < 1e-6
        assert np.abs(nt.rotation - np.array(m["rotation"])).max() < 1e-6
    first = nts[0]
    assert first.unit_id() == "1GID|1|A|G|103"
    assert np.allclose(first.base_center, [52.65518173, 52.42990909, 90.244], atol=1e-8)


def test_kabsch_known_answers(golden):
    """tests/superpositions_test.py:11-119 (octahedron 90 deg about z; 45 deg rotations have
    angle_of_rotation == pi/4) and the reference's outputs on seeded problems."""
    geo = golden("geometry.json.gz")
    by = {k["name"]: k for k in geo["kabsch"]}
    U = O.besttransformation(by["octahedron_90z"]["set1"], by["octahedron_90z"]["set2"])[0]
    np.testing.assert_almost_equal(U, [[0, 1, 0], [-1, 0, 0], [0, 0, 1]], decimal=7)
    for ax in "xyz":
        U = O.besttransformation(by["p4_45" + ax]["set1"], by["p4_45" + ax]["set2"])[0]
        np.testing.assert_almost_equal(O.angle_of_rotation(U), np.pi / 4, decimal=7)
    for k in geo["kabsch"]:
        U, new1, mean1, rmsd, sse, mean2 = O.besttransformation(k["set1"], k["set2"])
        assert np.abs(U - np.array(k["U"])).max() < 1e-12
        assert np.abs(new1 - np.array(k["new1"])).max() < 1e-11
        assert abs(sse - k["sse"]) < 1e-10 and abs(rmsd - k["rmsd"]) < 1e-12
        Uw, _, _, rmsdw, ssew = O.besttransformation_weighted(k["set1"], k["set2"], k["weights"])
        assert np.abs(Uw - np.array(k["Uw"])).max() < 1e-12 and abs(ssew - k["ssew"]) < 1e-10


def test_discrepancy_known_answers(golden):
    """tests/discrepancy_tests.py:301-335: 0.9402 and 1.1414 to 3 decimals; assertion behaviour
    :337-348; nucleotide_superposition_test.py:49-77 sarcin values."""
    geo = golden("geometry.json.gz")
    by = {k["name"]: k for k in geo["discrepancy_kats"]}

    def args(k):
        return ([np.array(c) for c in k["centers1"]], [np.array(r) for r in k["rotations1"]],
                [np.array(c) for c in k["centers2"]], [np.array(r) for r in k["rotations2"]])

    np.testing.assert_almost_equal(0.9402, O.matrix_discrepancy(*args(by["simple"])), decimal=3)
    np.testing.assert_almost_equal(1.1414, O.matrix_discrepancy(*args(by["another"])), decimal=3)
    for k in geo["discrepancy_kats"]:
        assert abs(O.matrix_discrepancy(*args(k)) - k["d"]) < 1e-13
    assert abs(by["sarcin7"]["d"] - 0.1159) < 1e-3
    k = by["simple"]
    assert O.matrix_discrepancy_cutoff(*args(k), 2.5) == pytest.approx(k["cutoff_2.5"], abs=1e-13)
    assert O.matrix_discrepancy_cutoff(*args(k), 0.5) is None and k["cutoff_0.5"] is None
    with pytest.raises(AssertionError):
        O.matrix_discrepancy([], [], [], [])
    a = list(args(by["another"]))
    a[2] = a[2][:-1]
    a[3] = a[3][:-1]
    with pytest.raises(AssertionError):
        O.matrix_discrepancy(*a)


def test_all_vs_all_block(golden):
    av = golden("geometry.json.gz")["all_vs_all"]
    C, R = np.array(av["centers"]), np.array(av["rotations"])
    sub = 12
    D = O.all_vs_all_discrepancy(C[:sub], R[:sub])
    assert np.abs(D - np.array(av["D"])[:sub, :sub]).max() < 1e-13
    has = np.ones((sub, av["n"]), dtype=bool)
    has[:, 0] = False
    Dn = O.all_vs_all_discrepancy(C[:sub], R[:sub], has)
    assert np.abs(Dn - np.array(av["D_norot0"])[:sub, :sub]).max() < 1e-13


def test_modified_residues_match_reference(golden):
    """Modified nucleotides (PSU, 5MC, OMG, 2MG, 1MA, MA6, H2U): parent-atom mapping, centre =
    meanR - U.meanS, and the reference's insertion of ideal-position copies of every mapped base atom when
    no hydrogen is observed (components.py:505-513)."""
    spec = golden("spec_1gid_modified.json.gz")
    frames = golden("frames_1gid_modified.json.gz")
    g = golden("annot_1gid_modified_all9.json.gz")
    nts = O.build_nts(spec)
    assert sum(1 for nt in nts if nt.sequence in O.MODIFIED) == 90
    for nt, f in zip(nts, frames):
        assert nt.unit_id() == f["unit_id"]
        assert np.abs(nt.base_center - np.array(f["center"])).max() == 0.0
        assert np.abs(nt.rotation - np.array(f["rotation"])).max() == 0.0
        for n, x in f["centers"].items():                       # AtomProxy means of observed + ideal
            assert np.abs(nt.center(n) - np.array(x)).max() == 0.0
        mine = sorted((n, tuple(x)) for n, x in zip(nt.atom_names, nt.atom_xyz) if n in nt.base_atom_names())
        ref = sorted((b[0], tuple(b[1:])) for b in f["base_points"])
        assert [m[0] for m in mine] == [r[0] for r in ref]
        assert np.abs(np.array([m[1] for m in mine]) - np.array([r[1] for r in ref])).max() == 0.0
    out, c2i, tr = O.annotate(spec, g["categories"], nts)
    assert tr["candidates"] == g["traces"]["candidates"]
    for key in ("sO", "stacking", "sugar_ribose", "backbone"):
        assert tr[key] == g["traces"][key], key
    _close_rows(tr["coplanar"], g["traces"]["coplanar"])
    _close_rows(tr["basepair"], g["traces"]["basepair"])
    assert {k: [list(t) for t in v] for k, v in out.items()} == g["interaction_to_list_of_tuples"]


@pytest.mark.parametrize("case", ["detail", "focus_cww_ths_stacking", "focus_cww_coplanar_covalent"])
def test_category_subsets_and_focused_interactions(golden, case):
    """Category subsets and focused interaction lists (focus_basepair_cutoffs, NA:88-136) against the reference."""
    spec = golden("spec_1gid_full.json.gz")
    g = golden("annot_1gid_full_category_subsets.json.gz")[case]
    out, c2i, tr = O.annotate(spec, g["categories"])
    got = {k: [list(t) for t in v] for k, v in out.items()}
    assert got == g["interaction_to_list_of_tuples"]
    assert {k: sorted(v) for k, v in c2i.items()} == g["category_to_interactions"]
    assert "Found %d nucleotide" % tr["count_pair"] in g["found_line"]


def _inflated_cutoffs(only_largest):
    """Same table as tools/make_golden.py::inflated_cutoffs, built from the oracle's copy of the shipped tables."""
    import copy
    focused = copy.deepcopy(O.focus_basepair_cutoffs(O.nt_nt_cutoffs(), []))
    sizes = {(c, sg): sum(len(v) for v in focused[c][sg].values()) for c in focused for sg in focused[c]}
    largest = max(sizes, key=sizes.get)
    for c in focused:
        for sg in focused[c]:
            if only_largest and (c, sg) != largest:
                continue
            for inter, subs in focused[c][sg].items():
                for sub in list(subs.keys()):
                    box = dict(subs[sub])
                    for k in ("xmin", "xmax", "ymin", "ymax"):
                        box[k] = box[k] + 0.35
                    box["gapmax"] = box.get("gapmax", 0.0) + 0.2
                    subs[sub + 50] = box
    return focused


@pytest.mark.parametrize("only_largest", [True, False])
def test_caller_supplied_cutoff_table(golden, only_largest):
    """focused_basepair_cutoffs passed by the caller (third argument of annotate_nt_nt_in_structure, NA:1213)."""
    spec = golden("spec_1gid_full.json.gz")
    g = golden("annot_1gid_full_custom_cutoffs.json.gz")["only_largest" if only_largest else "all"]
    categories = {k: [] for k in ("basepair", "basepair_detail", "stacking", "sO", "backbone", "coplanar",
                                  "sugar_ribose", "covalent", "near")}
    out, c2i, _ = O.annotate(spec, categories, focused=_inflated_cutoffs(only_largest))
    assert {k: [list(t) for t in v] for k, v in out.items()} == g["interaction_to_list_of_tuples"]
    assert {k: sorted(v) for k, v in c2i.items()} == g["category_to_interactions"]


@pytest.mark.parametrize("name", ["1gid_full", "1gid_modified", "1gid_variants", "1s72_strand", "chi_sweep"])
def test_bond_orientation_matches_reference(golden, name):
    """annotate_bond_orientation (NA_unit_annotation.py:48-171): the unmodified reference function's rows
    (tools/make_golden.py glycosidic) on the fixtures and on a 15-degree sweep of the glycosidic torsion."""
    g = golden("glycosidic.json.gz")[name]
    rows = O.bond_orientation(golden("spec_%s.json.gz" % name))
    assert [{k: r[k] for k in ("unit_id", "orientation", "chi_degree")} for r in rows] == g["rows"]
    if name == "chi_sweep":
        assert {r["orientation"] for r in rows} == {"anti", "syn", "int_syn"}


def _search_case(case):
    centers = np.array([[np.nan if v is None else v for v in c] for c in case["centers"]], dtype=float)
    return centers, case["models"]


@pytest.mark.parametrize("name", ["1gid_full", "1gid_variants"])
def test_search_screen_matches_reference(golden, name):
    """fixed_radius_search / get_pairlist (fr3d/search/pair_processing.py), imported unmodified by
    tools/make_golden.py search: same triples (orientation, order, squared distances bit for bit), same slices."""
    case = golden("search_screen.json.gz")[name]
    centers, models = _search_case(case)
    for cutoff, want in case["searches"].items():
        got = O.fixed_radius_search(len(centers), models, centers, float(cutoff))
        assert [[a, b, c] for a, b, c in got] == want
    pl = O.get_pairlist(case["Q"], models, centers)
    got = {"%d,%d" % (i, j): [list(p) for p in v] for i, d in pl.items() for j, v in d.items()}
    assert got == case["pairlist"]
