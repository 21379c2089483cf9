"""Parity of the CUDA path (through the public API / C ABI) against the reference golden vectors
and the oracle.  Integer / label work must be identical; fp64 outputs within 1e-9 (the tolerance
BASELINE.json states)."""

import copy

import numpy as np
import pytest

from conftest import ALL9
from fr3d_python_b200 import _lib, packing, synthetic
from fr3d_python_b200 import definitions as defs
from fr3d_python_b200.classifiers import NA_pairwise_interactions as NA
from fr3d_python_b200.data import Structure
from fr3d_python_b200.engine import Engine
from fr3d_python_b200.geometry import frames as gframes
from oracle import fr3d_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-9


def oracle_tuples(spec, categories, focused=None):
    nts = O.build_nts(spec)
    out, c2i, tr = O.annotate(spec, categories, nts, focused=focused)
    res = {k: [(nts[a].unit_id(), nts[b].unit_id(), c) for a, b, c in v] for k, v in out.items() if v}
    return res, {k: set(v) for k, v in c2i.items() if v}, tr, nts


def as_plain(result):
    i2t, c2i = result[0], result[1]
    return {k: list(v) for k, v in i2t.items() if v}, {k: set(v) for k, v in c2i.items() if v}


# ------------------------------------------------------------------ frames (K1)
@pytest.mark.parametrize("name", ["1gid_full", "1s72_strand", "1s72_sarcin", "1gid_variants"])
def test_frames_and_hydrogens_match_reference(golden, name):
    spec = golden("spec_%s.json.gz" % name)
    want = golden("frames_%s.json.gz" % name)
    b = packing.pack_specs(spec)
    status = gframes.fit_batch_frames(b, place_h=True)
    assert np.all(status == 0)
    for k, f in enumerate(want):
        assert np.abs(b.center[k] - np.array(f["center"])).max() < TOL
        assert np.abs(b.rot[k].reshape(3, 3) - np.array(f["rotation"])).max() < TOL
        parent = defs.PARENTS[b.meta[k, 0]]
        for hname, xyz in f["hydrogens"].items():
            slot = defs.SLOT_OF[parent][hname]
            assert b.base_mask[k] >> slot & 1
            assert np.abs(b.base_xyz[k, slot] - np.array(xyz)).max() < TOL
        assert b.base_mask[k] & _lib.MASK_HAS_FRAME


def test_frames_match_matlab_golden(golden):
    b = packing.pack_specs(golden("spec_1gid_base.json.gz"))
    gframes.fit_batch_frames(b)
    for k, m in enumerate(golden("matlab_1gid_frames.json.gz")):
        assert np.abs(b.center[k] - np.array(m["center"])).max() < 1e-6       # inputs carry 6 decimals
        assert np.abs(b.rot[k].reshape(3, 3) - np.array(m["rotation"])).max() < 1e-6


def test_component_api_fits_frames_on_gpu(golden):
    spec = golden("spec_1s72_strand.json.gz")
    st = Structure.from_spec(spec)
    want = golden("frames_1s72_strand.json.gz")
    res = st.residues(type=["RNA linking", "DNA linking"])
    assert len(res) == 10
    for c, f in zip(res, want):
        assert c.unit_id() == f["unit_id"]
        assert isinstance(c.rotation_matrix, np.matrix)
        assert np.abs(np.asarray(c.rotation_matrix) - np.array(f["rotation"])).max() < TOL
        assert np.abs(c.centers["base"] - np.array(f["center"])).max() < TOL
        for hname, xyz in f["hydrogens"].items():
            assert np.abs(c.centers[hname] - np.array(xyz)).max() < TOL
        gly = "N9" if c.sequence in "AG" else "N1"
        assert np.array_equal(c.centers["glycosidic"], c.centers[gly])


def test_degenerate_and_unknown_residues(golden):
    spec = copy.deepcopy(golden("spec_1s72_strand.json.gz"))
    spec["nts"][2]["atoms"] = [a for a in spec["nts"][2]["atoms"] if a[0] in ("N1", "C2", "P", "O2'")]   # 2 base atoms
    spec["nts"][4]["sequence"] = "XYZ"                                                                   # unknown residue
    b = packing.pack_specs(spec)
    status = gframes.fit_batch_frames(b)
    assert status[2] == -2 and status[4] == -1 and np.all(np.delete(status, [2, 4]) == 0)
    assert not (b.base_mask[2] & _lib.MASK_HAS_FRAME) and not (b.base_mask[4] & _lib.MASK_HAS_FRAME)
    res = NA.annotate_batch(spec, ALL9)[0]
    pairs = [t for v in res[0].values() for t in v]
    bad = {b.unit_ids()[2], b.unit_ids()[4]}
    assert not any((t[0] in bad or t[1] in bad) and t[2] is not None for t in pairs)     # only p_ links mention them


# ------------------------------------------------------------------ search (K2/K3)
def test_candidate_pairs_in_reference_order(golden):
    spec = golden("spec_1gid_full.json.gz")
    g = golden("annot_1gid_full_all9.json.gz")
    results, ex = NA.annotate_batch(spec, ALL9, return_raw=True)
    pi, pj, pr = ex["pairs"]
    order = np.lexsort((pj, pi, pr))
    got = [[int(a), int(b)] for a, b in zip(pi[order], pj[order])]
    assert got == g["traces"]["candidates"]
    assert ex["n_pairs"] == 1983


def test_search_edge_cases(golden):
    eng = Engine.get()
    assert as_plain(NA.annotate_batch({"pdb": "E", "nts": []}, ALL9)[0]) == ({}, {})      # empty structure
    spec = copy.deepcopy(golden("spec_1s72_strand.json.gz"))
    # symmetry copy of nt 3 a few Angstrom away, same chain and index: the reference visits the
    # pair in both orders (NA:684-688); negative coordinates exercise floor()
    for nt in spec["nts"]:
        for a in nt["atoms"]:
            a[1] -= 200.0
            a[2] -= 160.0
    twin = copy.deepcopy(spec["nts"][3])
    twin["symmetry"] = "2_555"
    for a in twin["atoms"]:
        a[3] += 3.4
    spec["nts"].append(twin)
    # alt-id twins of nt 6: never compared with each other (NA:706-713)
    alt = copy.deepcopy(spec["nts"][6])
    spec["nts"][6]["alt_id"] = "A"
    alt["alt_id"] = "B"
    for a in alt["atoms"]:
        a[1] += 2.5
    spec["nts"].append(alt)
    want, wc, tr, nts = oracle_tuples(spec, ALL9)
    results, ex = NA.annotate_batch(spec, ALL9, return_raw=True)
    pi, pj, pr = ex["pairs"]
    order = np.lexsort((pj, pi, pr))
    got = [(int(a), int(b)) for a, b in zip(pi[order], pj[order])]
    assert got == [tuple(p) for p in tr["candidates"]]
    assert (3, 10) in got or (10, 3) in got                      # the symmetry twin is a partner
    assert (6, 11) not in got and (11, 6) not in got
    assert as_plain(results[0]) == (want, wc)
    del eng


def test_out_of_range_coordinates_raise(golden):
    spec = copy.deepcopy(golden("spec_1s72_strand.json.gz"))
    for a in spec["nts"][0]["atoms"]:
        a[1] += 1.0e6
    with pytest.raises(_lib.Fr3dError):
        NA.annotate_batch(spec, ALL9)


def test_capacity_overflow_is_retried_not_truncated(golden):
    spec = golden("spec_1gid_full.json.gz")
    batch = packing.pack_specs(spec)
    eng = Engine.get()
    bt, atoms = NA._box_table(NA._default_focused(ALL9['basepair']), NA.load_ideal_basepair_hydrogen_bonds())
    tiny = eng.make_plan(batch.n, pair_cap=100, hit_cap=10)
    hits, n_pairs, ex = eng.annotate_batch(batch, bt, NA.category_mask(ALL9), plan=tiny)
    assert n_pairs == 1983 and len(hits) == 783


# ------------------------------------------------------------------ full annotation (K1-K4 + host tail)
@pytest.mark.parametrize("spec_name,annot_name", [
    ("spec_1s72_strand.json.gz", "annot_1s72_strand_all9.json.gz"),
    ("spec_1gid_full.json.gz", "annot_1gid_full_all9.json.gz"),
    ("spec_1gid_base.json.gz", "annot_1gid_base_bp_stack_cp.json.gz"),
    ("spec_1gid_base.json.gz", "annot_1gid_base_bp.json.gz"),
    # DNA, missing atoms, alt-id twins, insertion codes, a second model and a symmetry copy (tools/make_golden.py variants)
    ("spec_1gid_variants.json.gz", "annot_1gid_variants_all9.json.gz"),
])
def test_annotation_identical_to_reference(golden, spec_name, annot_name):
    spec = golden(spec_name)
    g = golden(annot_name)
    uid = g["unit_ids"]
    want = {k: [(uid[a], uid[b], c) for a, b, c in v] for k, v in g["interaction_to_list_of_tuples"].items()}
    wantc = {k: set(v) for k, v in g["category_to_interactions"].items()}
    # batched entry point, frames fitted on the GPU
    res = NA.annotate_batch(spec, g["categories"])[0]
    i2t, c2i = as_plain(res)
    assert i2t == want and c2i == wantc                 # same rows, same order as fr3d-python
    assert list(res[0].keys()) == [k for k, v in g["interaction_to_list_of_tuples"].items() if v]   # same dict order
    # the host form of the tail (postprocess.py) gives the same
    assert as_plain(NA.annotate_batch(spec, g["categories"], tail="host")[0]) == (want, wantc)
    # reference entry point on Structure / Component objects
    st = Structure.from_spec(spec)
    r = NA.annotate_nt_nt_in_structure(st, g["categories"])
    assert len(r) == 4
    i2t2, c2i2 = as_plain(r)
    assert i2t2 == want and c2i2 == wantc
    assert isinstance(r[2], dict) and "allStates" in r[2] and r[3] == {}


def test_get_datapoint_is_refused(golden):
    st = Structure.from_spec(golden("spec_1s72_strand.json.gz"))
    with pytest.raises(NotImplementedError):
        NA.annotate_nt_nt_in_structure(st, ALL9, get_datapoint=True)


def test_chain_filter(golden):
    spec = golden("spec_1gid_full.json.gz")
    st = Structure.from_spec(spec)
    r = NA.annotate_nt_nt_in_structure(st, {'basepair': NA.Leontis_Westhof_basepairs, 'stacking': []}, chains=["A"])
    sub = {"pdb": spec["pdb"], "nts": [nt for nt in spec["nts"] if nt["chain"] == "A"]}
    want, wc, _, _ = oracle_tuples(sub, {'basepair': NA.Leontis_Westhof_basepairs, 'stacking': []})
    assert as_plain(r) == (want, wc)


def test_perturbed_batch_against_oracle():
    """Seeded perturbed copies annotated in ONE batch launch == oracle per structure."""
    base = synthetic.pack_base()
    batch = synthetic.make_structure_batch(base, 6, seed0=4000)
    results = NA.annotate_batch(batch, ALL9)
    for s in range(6):
        want, wc, _, _ = oracle_tuples(synthetic.batch_to_spec(batch, s), ALL9)
        assert as_plain(results[s]) == (want, wc), "structure %d" % s


def test_dna_and_missing_atoms_against_oracle(golden):
    spec = copy.deepcopy(golden("spec_1gid_full.json.gz"))
    nts = O.build_nts(spec)
    std = defs.NAbasecoordinates["DT"]
    for k, nt in enumerate(spec["nts"][:120]):
        if k % 3 == 0 and nt["sequence"] in ("A", "C", "G"):
            nt["sequence"] = "D" + nt["sequence"]
            nt["atoms"] = [a for a in nt["atoms"] if a[0] != "O2'"]
        elif k % 3 == 1 and nt["sequence"] == "U":
            nt["sequence"] = "DT"
            c7 = nts[k].base_center + nts[k].rotation.dot(np.array(std["C7"]))
            nt["atoms"] = [a for a in nt["atoms"] if a[0] != "O2'"] + [["C7", float(c7[0]), float(c7[1]), float(c7[2])]]
        elif k % 7 == 2:
            nt["atoms"] = [a for a in nt["atoms"] if a[0] not in ("P", "OP1", "O4'", "C2'")]
        elif k % 11 == 5:
            nt["atoms"] = [a for a in nt["atoms"] if a[0] not in ("N9", "N1")]            # no glycosidic atom
        elif k % 13 == 6:
            nt["atoms"] = [a for a in nt["atoms"] if a[0] not in ("N3", "O2", "C5")]
    want, wc, tr, _ = oracle_tuples(spec, ALL9)
    got = as_plain(NA.annotate_batch(spec, ALL9)[0])
    assert got == (want, wc)
    assert any("DT" in t[0] or "DT" in t[1] for v in want.values() for t in v)


def _inflated_cutoffs(only_largest):
    """focused_basepair_cutoffs with every subcategory box duplicated under a new subcategory key with shifted
    limits: more than 16 boxes per (combination, sign) entry and - unless only the largest entry is inflated -
    more than 256 boxes in total, which are the two table sizes the screen kernel treats differently."""
    focused = copy.deepcopy(NA.focus_basepair_cutoffs(NA.nt_nt_cutoffs, []))
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
def test_custom_cutoff_tables_against_oracle(golden, only_largest):
    """Caller-supplied focused_basepair_cutoffs (the reference's third argument) with large tables."""
    from fr3d_python_b200.engine import BoxTable
    focused = _inflated_cutoffs(only_largest)
    bt = BoxTable(focused)
    assert bt.index[:, :, 1].max() > 16
    assert (len(bt.boxes) > 256) == (not only_largest)
    spec = golden("spec_1gid_full.json.gz")
    want, wc, _, _ = oracle_tuples(spec, ALL9, focused=focused)
    got = as_plain(NA.annotate_batch(spec, ALL9, focused_basepair_cutoffs=focused)[0])
    assert got == (want, wc)
    # ... and against the rows the unmodified reference produced with the same table (tools/make_golden.py cutoffs)
    gg = golden("annot_1gid_full_custom_cutoffs.json.gz")
    g, uid = gg["only_largest" if only_largest else "all"], gg["unit_ids"]
    ref = {k: [(uid[a], uid[b], c) for a, b, c in v] for k, v in g["interaction_to_list_of_tuples"].items() if v}
    assert got[0] == ref
    assert got[1] == {k: set(v) for k, v in g["category_to_interactions"].items() if v}
    if not only_largest:
        base_want, _, _, _ = oracle_tuples(spec, ALL9)
        assert want != base_want                   # the inflated table really changes the annotation


def test_observed_hydrogens_are_kept(golden):
    """High-resolution input with observed (and mislabelled) amino hydrogens, components.py:405-445."""
    spec = copy.deepcopy(golden("spec_1s72_strand.json.gz"))
    fr = golden("frames_1s72_strand.json.gz")
    for k in (1, 3):                         # A and G: put H61/H62 (H21/H22) in, swapped and shifted
        h = fr[k]["hydrogens"]
        names = ("H61", "H62") if "H61" in h else ("H21", "H22")
        spec["nts"][k]["atoms"] += [[names[0]] + [v + 0.05 for v in h[names[1]]], [names[1]] + [v + 0.05 for v in h[names[0]]]]
    b = packing.pack_specs(spec)
    gframes.fit_batch_frames(b)
    for k in (1, 3):
        parent = defs.PARENTS[b.meta[k, 0]]
        h = fr[k]["hydrogens"]
        names = ("H61", "H62") if "H61" in h else ("H21", "H22")
        for n in names:      # relabelled back, observed coordinates kept (not replaced by ideal ones)
            assert np.abs(b.base_xyz[k, defs.SLOT_OF[parent][n]] - (np.array(h[n]) + 0.05)).max() < 1e-12


def test_host_tail_in_worker_processes_gives_the_same_results():
    """annotate_batch(tail_workers=n): the per-structure host tail in forked worker processes."""
    batch = synthetic.make_structure_batch(synthetic.pack_base(), 40, seed0=321)
    one = NA.annotate_batch(batch, ALL9, tail_workers=1, tail="host")
    many = NA.annotate_batch(batch, ALL9, tail_workers=4, tail="host")
    assert len(one) == len(many) == 40
    for a, b in zip(one, many):
        assert list(a[0].keys()) == list(b[0].keys())            # same labels in the same order ...
        assert as_plain(a) == as_plain(b)                        # ... with the same rows in the same order


def test_atom_far_from_its_base_goes_through_fp64(golden):
    """An observed hydrogen hundreds of Angstrom from its base (garbage input) is outside the error bounds of the
    single-precision kernels: such a record is flagged (classify_screen32.cuh, S32_FAR) and its stacking pairs are
    decided by the fp64 kernel.  Whatever the path, the result must be the oracle's."""
    spec = copy.deepcopy(golden("spec_1gid_full.json.gz"))
    moved = 0
    for k, nt in enumerate(spec["nts"]):
        if k % 17 == 3:
            name = {"A": "H8", "G": "H8", "C": "H6", "U": "H6"}[nt["sequence"]]
            heavy = np.array([a[1:] for a in nt["atoms"] if a[0] in ("C2", "C4", "C6")]).mean(axis=0)
            nt["atoms"] = [a for a in nt["atoms"] if a[0] != name] + [[name] + list(heavy + np.array([0.0, 700.0, 0.004]))]
            moved += 1
    assert moved > 10
    want, wc, _, _ = oracle_tuples(spec, ALL9)
    got = as_plain(NA.annotate_batch(spec, ALL9)[0])
    assert got == (want, wc)


def test_s1s72_like_structure_against_oracle():
    """C2 stand-in: 9 rotated, noised copies of 1GID-full in ONE structure, 2 844 nt."""
    base = synthetic.pack_base()
    t = synthetic.make_tiled_structure(base, 9, (3, 3, 1), seed=1072, pdb="1S72L")
    res, ex = NA.annotate_batch(t, ALL9, return_raw=True)
    want, wc, tr, _ = oracle_tuples(synthetic.batch_to_spec(t, 0), ALL9)
    assert ex["n_pairs"] == len(tr["candidates"])
    assert as_plain(res[0]) == (want, wc)


def test_full_size_batch_properties():
    """C4 (1000 structures) and C3 (25 280 nt in one structure): size-independent properties."""
    base = synthetic.pack_base()
    batch = synthetic.make_structure_batch(base, 1000, seed0=4000)
    eng = Engine.get()
    bt, atoms = NA._box_table(NA._default_focused(ALL9['basepair']), NA.load_ideal_basepair_hydrogen_bonds())
    mask = NA.category_mask(ALL9)
    hits, n_pairs, ex = eng.annotate_batch(batch, bt, mask, return_frames=True)
    hits2, n_pairs2, _ = eng.annotate_batch(batch, bt, mask)
    key = lambda h: np.lexsort((h['slot'], h['j'], h['i'], h['rank']))
    assert n_pairs == n_pairs2 and np.array_equal(hits[key(hits)], hits2[key(hits2)])     # deterministic set
    pi, pj, pr = ex["pairs"]
    sid = np.searchsorted(batch.struct_off, pi, side='right') - 1
    assert np.array_equal(sid, np.searchsorted(batch.struct_off, pj, side='right') - 1)   # never across structures
    d = np.linalg.norm(ex["center"][pi] - ex["center"][pj], axis=1)
    assert d.min() >= 2.0 and d.max() <= 12.0
    code = pi.astype(np.int64) * batch.n + pj
    assert len(np.unique(code)) == len(code)                                             # each oriented pair once
    und = np.minimum(pi, pj).astype(np.int64) * batch.n + np.maximum(pi, pj)
    assert len(np.unique(und)) == len(und)                                               # and in one orientation only
    # every unordered pair within 12 A is found: brute force on three structures
    for s in (0, 499, 999):
        lo, hi = batch.struct_off[s], batch.struct_off[s + 1]
        c = ex["center"][lo:hi]
        dd = np.linalg.norm(c[:, None] - c[None], axis=2)
        want = int(((dd >= 2.0) & (dd <= 12.0)).sum() // 2)
        assert int((sid == s).sum()) == want
    # rotations are proper and orthonormal
    R = ex["rot"].reshape(-1, 3, 3)
    assert np.abs(np.einsum('nij,nkj->nik', R, R) - np.eye(3)).max() < 1e-12
    assert np.abs(np.linalg.det(R) - 1).max() < 1e-12
    # structure k of the batch == the same structure annotated alone
    res_all = NA.annotate_batch(packing.select_structures(batch, [0, 499, 999]), ALL9)
    sub = postprocess_subset(batch, hits, bt, atoms, [0, 499, 999])
    for a, b in zip(res_all, sub):
        assert as_plain(a) == as_plain(b)
    # C3: one 25 280-nt structure
    big = synthetic.make_tiled_structure(base, 80, (5, 4, 4), seed=25000, pdb="C3")
    hb, npb, exb = eng.annotate_batch(big, bt, mask, return_frames=True)
    pi, pj, pr = exb["pairs"]
    d = np.linalg.norm(exb["center"][pi] - exb["center"][pj], axis=1)
    assert big.n == 25280 and d.min() >= 2.0 and d.max() <= 12.0
    und = np.minimum(pi, pj).astype(np.int64) * big.n + np.maximum(pi, pj)
    assert len(np.unique(und)) == len(und)
    assert np.all(hb['slot'] <= 8)


# ------------------------------------------------------------------ device tail (N1)
def _rows_by_structure(res, s):
    r = res.rows_of(s)
    return [(res.labels.name(l), a, b, c) for l, a, b, c in zip(r['label'].tolist(), r['nt1'].tolist(), r['nt2'].tolist(),
                                                                 r['crossing'].tolist())]


def test_device_tail_equals_host_tail_and_oracle():
    """fr3d_annotation_tail (same-edge cleanup, emission, crossing numbers, covalent links on the GPU) against the
    host form of the tail on 64 structures at three noise levels - rows, row order AND dict key order - and
    against the oracle on 16 of them."""
    base = synthetic.pack_base()
    for sigma, seed0, n_oracle in ((0.10, 4000, 6), (0.50, 7000, 5), (1.0, 9000, 5)):
        batch = synthetic.make_structure_batch(base, 64 if sigma == 0.10 else 24, seed0=seed0, sigma=sigma)
        dev = NA.annotate_batch(batch, ALL9)
        host = NA.annotate_batch(batch, ALL9, tail="host")
        assert len(dev) == len(host) == batch.n_structures
        for s in range(batch.n_structures):
            assert list(dev[s][0].keys()) == list(host[s][0].keys()), "structure %d, sigma %g" % (s, sigma)
            assert as_plain(dev[s]) == as_plain(host[s]), "structure %d, sigma %g" % (s, sigma)
        rng = np.random.default_rng(seed0)
        for s in rng.choice(batch.n_structures, n_oracle, replace=False).tolist():
            want, wc, _, _ = oracle_tuples(synthetic.batch_to_spec(batch, s), ALL9)
            assert as_plain(dev[s]) == (want, wc), "structure %d, sigma %g" % (s, sigma)


def test_full_size_batch_random_structures_against_oracle():
    """C4 at bench size (1000 structures in one pass, device tail): 16 random structures against the oracle, row
    arrays consistent, every structure's block in bounds."""
    base = synthetic.pack_base()
    batch = synthetic.make_structure_batch(base, 1000, seed0=4000)
    res = NA.annotate_batch(batch, ALL9, as_arrays=True)
    assert len(res) == 1000 and res.n_rows == len(res.rows)
    ends = np.sort(res.row_start.astype(np.int64) + res.row_count)
    starts = np.sort(res.row_start.astype(np.int64))
    assert starts[0] == 0 and np.array_equal(starts[1:], ends[:-1]) and ends[-1] == len(res.rows)     # a partition
    sid1 = np.searchsorted(batch.struct_off, res.rows['nt1'], side='right') - 1
    sid2 = np.searchsorted(batch.struct_off, res.rows['nt2'], side='right') - 1
    assert np.array_equal(sid1, sid2)
    for s in np.random.default_rng(5).choice(1000, 16, replace=False).tolist():
        assert np.all(sid1[res.row_start[s]:res.row_start[s] + res.row_count[s]] == s)
        want, wc, _, _ = oracle_tuples(synthetic.batch_to_spec(batch, s), ALL9)
        assert as_plain(res[s]) == (want, wc), "structure %d" % s


def test_device_tail_row_order_is_the_append_order(golden):
    """The flat row array lists the tuples in the order the reference appends them: grouping it by label (stable)
    gives the golden lists, and the first occurrences give the golden dict order."""
    spec = golden("spec_1gid_variants.json.gz")
    g = golden("annot_1gid_variants_all9.json.gz")
    res = NA.annotate_batch(spec, g["categories"], as_arrays=True)
    rows = _rows_by_structure(res, 0)
    grouped = {}
    for l, a, b, c in rows:
        grouped.setdefault(l, []).append([a, b, None if c < 0 else c])
    want = {k: v for k, v in g["interaction_to_list_of_tuples"].items() if v}
    assert list(grouped.keys()) == list(want.keys())
    assert grouped == want


def test_device_tail_in_the_pipeline_and_capacity_reruns():
    """engine.Pipeline with the device tail: rows of every chunking equal the single pass; a pipeline built with
    far too small capacities re-runs its chunks instead of raising or truncating."""
    import torch
    from fr3d_python_b200.engine import Pipeline, pin_batch
    base = synthetic.pack_base()
    batch = synthetic.make_structure_batch(base, 37, seed0=900)
    eng = Engine.get()
    bt, atoms, labels = NA._tables(NA._default_focused(ALL9['basepair']), NA.load_ideal_basepair_hydrogen_bonds())
    mask = NA.category_mask(ALL9)
    ref = NA.annotate_batch(batch, ALL9, as_arrays=True)
    pinned = pin_batch(torch, batch)
    n_hits = len(eng.annotate_batch(batch, bt, mask)[0])
    # (hit_cap = the exact hit count: the tail's one-pass binning by proportional shares overflows for every structure
    # with more hits than the average and the exact scan + scatter path takes over)
    for chunks, caps in ((1, None), (5, None), (37, None), (3, (100, 100, 50)), (1, (ref.n_pairs, n_hits, ref.n_rows)),
                         (2, (ref.n_pairs, n_hits, ref.n_rows))):
        kw = {} if caps is None else {"pair_cap": caps[0], "hit_cap": caps[1], "row_cap": caps[2]}
        pipe = Pipeline(eng, batch, pinned, bt, mask, n_chunks=chunks, labels=labels, **kw)
        for _ in range(2):
            out = pipe.run()
        assert out.n_pairs == ref.n_pairs and out.hits is None
        for s in range(batch.n_structures):
            got = out.rows[out.row_start[s]:out.row_start[s] + out.row_count[s]]
            assert np.array_equal(got, ref.rows_of(s)), "chunks %d structure %d" % (chunks, s)
    # one CUDA graph per pass (what bench.py's e2e replays): same rows, repeatedly
    pipe = Pipeline(eng, batch, pinned, bt, mask, n_chunks=4, labels=labels)
    for _ in range(3):
        out = pipe.run_graph()
        assert out.n_pairs == ref.n_pairs and out.n_rows == ref.n_rows
        for s in range(batch.n_structures):
            got = out.rows[out.row_start[s]:out.row_start[s] + out.row_count[s]]
            assert np.array_equal(got, ref.rows_of(s)), "graph, structure %d" % s
    # the single-pass entry point re-runs too
    tiny = eng.make_plan(batch.n, pair_cap=100, hit_cap=10)
    hits, n_pairs, ex = eng.annotate_batch(batch, bt, mask, plan=tiny, labels=labels)
    assert n_pairs == ref.n_pairs and len(ex["rows"]) == ref.n_rows


def test_sharded_annotator_single_rank_equals_annotate_batch():
    """parallel.ShardedAnnotator (what bench.py runs under torchrun) with one rank: tail kernels write into the
    rank's send block, the block goes to the host in one copy and is indexed in place - same rows as annotate_batch;
    structures of different sizes exercise the LPT order and the chunk layout."""
    from fr3d_python_b200 import parallel
    base = synthetic.pack_base()
    small = packing.pack_specs(synthetic.load_golden_spec("spec_1s72_strand"))
    specs = []
    a = synthetic.make_structure_batch(base, 9, seed0=50)
    b = synthetic.make_structure_batch(small, 7, seed0=70, sigma=0.05)
    for k in range(9):
        specs.append(synthetic.batch_to_spec(a, k))
        if k < 7:
            specs.append(synthetic.batch_to_spec(b, k))
    for k, sp in enumerate(specs):
        sp["pdb"] = "M%02d" % k
    batch = packing.pack_specs(specs)
    ref = NA.annotate_batch(batch, ALL9, as_arrays=True)
    for chunks in (1, 3):
        sh = parallel.ShardedAnnotator(batch, ALL9, rank=0, world_size=1, n_chunks=chunks)
        for _ in range(2):
            got = sh.run()
        assert got.n_rows == ref.n_rows and got.n_pairs == ref.n_pairs
        for s in range(batch.n_structures):
            assert np.array_equal(got.rows_of(s), ref.rows_of(s)), "chunks %d structure %d" % (chunks, s)
        assert as_plain(got[3]) == as_plain(ref[3])
        again = sh.run(upload=False)                      # on the records already in HBM (bench.py's resident loop)
        assert all(np.array_equal(again.rows_of(s), ref.rows_of(s)) for s in range(batch.n_structures))
        for _ in range(2):                                # graph replays (captured above) keep giving the same rows
            assert all(np.array_equal(sh.run().rows_of(s), ref.rows_of(s)) for s in range(batch.n_structures))
    plain = parallel.ShardedAnnotator(batch, ALL9, rank=0, world_size=1, n_chunks=2, use_graph=False).run()
    assert all(np.array_equal(plain.rows_of(s), ref.rows_of(s)) for s in range(batch.n_structures))


def test_repeated_passes_over_resident_records_are_identical():
    """bench.py's resident loop: K1..K4 + tail run again and again over the same device records.  K1 reads the
    uploaded compact base plane and its observed-mask plane and writes the full records elsewhere, so every pass
    must reproduce the first one bit for bit (a K1 that took its own placed hydrogens for observed ones would not)."""
    import torch
    from fr3d_python_b200 import tail as tailmod
    from fr3d_python_b200.engine import DeviceBatch, DeviceIdent, HIT_DTYPE, pin_batch
    batch = synthetic.make_structure_batch(synthetic.pack_base(), 50, seed0=4000)
    eng = Engine.get()
    bt, atoms, labels = NA._tables(NA._default_focused(ALL9['basepair']), NA.load_ideal_basepair_hydrogen_bonds())
    mask = NA.category_mask(ALL9)
    pinned = pin_batch(torch, batch)
    assert pinned["base_xyz"].shape[1] == 11
    ident = DeviceIdent(torch, eng.device, tailmod.tail_ident(batch))
    plan = eng.make_plan(batch.n, tail=(ident.n_struct, ident.n_ends))
    db = DeviceBatch(torch, eng.device, batch, pinned)
    key = lambda h: np.lexsort((h['slot'], h['j'], h['i'], h['rank']))
    first = None
    for it in range(3):
        eng.run_plan(db, plan, bt, mask, fit_frames=True, tail=(ident, labels))
        n_pairs, n_hits = eng.read_counts(plan)
        h = plan.hits[:n_hits * 32].cpu().numpy().view(HIT_DTYPE)
        state = (n_pairs, h[key(h)].tobytes(), eng.read_tail_counts(plan), db.t["base_xyz"].cpu().numpy().tobytes(),
                 db.t["base_mask"].cpu().numpy().tobytes())
        if first is None:
            first = state
        assert state == first, "pass %d differs from the first" % it
    ref_hits, ref_pairs, _ = eng.annotate_batch(batch, bt, mask)
    assert ref_pairs == first[0] and ref_hits[key(ref_hits)].tobytes() == first[1]


def test_device_tail_edge_cases(golden):
    """Empty structures inside a batch, a structure without any hit, multi-model / symmetry / insertion-code
    identities (covalent links per model + operator + chain), and the ranges the device tail refuses."""
    strand = golden("spec_1s72_strand.json.gz")
    variants = golden("spec_1gid_variants.json.gz")
    far = copy.deepcopy(strand)
    far["pdb"] = "FAR"
    far["nts"] = far["nts"][:1]                                  # one nt: no pair, no covalent link
    specs = [{"pdb": "E0", "nts": []}, strand, far, {"pdb": "E1", "nts": []}, variants]
    dev = NA.annotate_batch(specs, ALL9)
    host = NA.annotate_batch(specs, ALL9, tail="host")
    assert len(dev) == 5
    for s in range(5):
        assert list(dev[s][0].keys()) == list(host[s][0].keys())
        assert as_plain(dev[s]) == as_plain(host[s])
    assert as_plain(dev[0]) == ({}, {}) and as_plain(dev[2]) == ({}, {})
    bad = copy.deepcopy(strand)
    bad["nts"][3]["index"] = -4
    with pytest.raises(_lib.Fr3dError):
        NA.annotate_batch(bad, ALL9)
    assert len(NA.annotate_batch(bad, ALL9, tail="host")) == 1      # the host tail has no such limit


def postprocess_subset(batch, hits, bt, atoms, which):
    from fr3d_python_b200 import postprocess
    hits = postprocess.sort_hits(hits)
    sidx = np.searchsorted(batch.struct_off, hits['i'], side='right') - 1
    return [postprocess.build_results(batch, s, hits[sidx == s], bt, atoms, ALL9) for s in which]


@pytest.mark.parametrize("observed_h", [False, True])
def test_pipelined_path_equals_single_pass(observed_h):
    """engine.Pipeline (chunked, multi-stream, what bench.py's e2e times) returns the same hit set; with and
    without observed hydrogens (pin_batch uploads only base slots 0..10 when none is observed)."""
    import torch
    from fr3d_python_b200.engine import HEAVY_SLOTS, Pipeline, pin_batch
    base = synthetic.pack_base()
    batch = synthetic.make_structure_batch(base, 37, seed0=900)
    if observed_h:                       # give every 5th G an observed H22 (slot 15) displaced from the ideal spot
        fitted = packing.select_structures(batch, range(batch.n_structures))
        gframes.fit_batch_frames(fitted)
        slot = defs.SLOT_OF["G"]["H22"]
        assert slot >= HEAVY_SLOTS
        gs = np.nonzero(batch.meta[:, 0] == defs.PARENTS.index("G"))[0][::5]
        batch.base_xyz[gs, slot] = fitted.base_xyz[gs, slot] + 0.07
        batch.base_mask[gs] |= np.uint32(1 << slot)
    eng = Engine.get()
    bt, atoms = NA._box_table(NA._default_focused(ALL9['basepair']), NA.load_ideal_basepair_hydrogen_bonds())
    mask = NA.category_mask(ALL9)
    hits, n_pairs, _ = eng.annotate_batch(batch, bt, mask)
    pinned = pin_batch(torch, batch)
    assert pinned["base_xyz"].shape[1] == (16 if observed_h else HEAVY_SLOTS)
    for chunks in (1, 5, 37):
        pipe = Pipeline(eng, batch, pinned, bt, mask, n_chunks=chunks, pair_cap=n_pairs, hit_cap=len(hits))
        for _ in range(2):
            res = pipe.run()
        h2, np2 = res.hits, res.n_pairs
        key = lambda h: np.lexsort((h['slot'], h['j'], h['i'], h['rank']))
        assert np2 == n_pairs and np.array_equal(hits[key(hits)], h2[key(h2)])


def test_modified_residues_identical_to_reference(golden):
    """90 modified nucleotides (PSU 5MC OMG 2MG 1MA MA6 H2U) in 1GID-full: frames, the averaged
    observed/ideal coordinates and the final annotation against the unmodified reference."""
    spec = golden("spec_1gid_modified.json.gz")
    frames = golden("frames_1gid_modified.json.gz")
    g = golden("annot_1gid_modified_all9.json.gz")
    b = packing.pack_specs(spec)
    status = gframes.fit_batch_frames(b)
    assert np.all(status == 0)
    n_dup = 0
    for k, f in enumerate(frames):
        assert np.abs(b.center[k] - np.array(f["center"])).max() < TOL
        assert np.abs(b.rot[k].reshape(3, 3) - np.array(f["rotation"])).max() < TOL
        if f["sequence"] in defs.modified_table():
            parent = defs.get_parent(f["sequence"])
            m2p = defs.modified_table()[f["sequence"]]["modified_atom_to_parent"]
            for name, xyz in f["centers"].items():
                slot = defs.SLOT_OF[parent][m2p[name]]
                assert b.base_mask[k] >> slot & 1
                assert np.abs(b.base_xyz[k, slot] - np.array(xyz)).max() < TOL
            assert bin(int(b.base_mask[k]) & 0xFFFF).count("1") == len(f["centers"])
            n_dup += bool(b.meta[k, 1] & _lib.FLAG_DUP)
    assert n_dup > 50
    uid = g["unit_ids"]
    want = {k: [(uid[a], uid[c], x) for a, c, x in v] for k, v in g["interaction_to_list_of_tuples"].items()}
    wantc = {k: set(v) for k, v in g["category_to_interactions"].items()}
    assert as_plain(NA.annotate_batch(spec, g["categories"])[0]) == (want, wantc)
    # through Component objects (ideal copies live in the atom lists, like in the reference)
    st = Structure.from_spec(spec)
    assert as_plain(NA.annotate_nt_nt_in_structure(st, g["categories"])) == (want, wantc)
    # a second K1 pass leaves residues holding averaged coordinates alone
    again = gframes.fit_batch_frames(b)
    assert np.all(again == 0)
    assert np.abs(b.center[0] - np.array(frames[0]["center"])).max() < TOL


@pytest.mark.parametrize("mode", ["0", "1"])
def test_alternate_classify_paths_give_the_same_hits(mode, tmp_path):
    """fr3d_classify_pairs without a workspace runs one warp-per-pair launch over all pairs (FR3D_CLASSIFY_MODE=0
    forces it), =1 one thread-per-pair launch: both must produce exactly the hit set of the default screened path.
    The mode is read once per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from conftest import ALL9\n"
        "from fr3d_python_b200 import synthetic\n"
        "from fr3d_python_b200.classifiers import NA_pairwise_interactions as NA\n"
        "from fr3d_python_b200.engine import Engine\n"
        "batch = synthetic.make_structure_batch(synthetic.pack_base(), 40, seed0=77)\n"
        "bt, _ = NA._box_table(NA._default_focused(ALL9['basepair']), NA.load_ideal_basepair_hydrogen_bonds())\n"
        "hits, n_pairs, _ = Engine.get().annotate_batch(batch, bt, NA.category_mask(ALL9))\n"
        "order = np.lexsort((hits['slot'], hits['j'], hits['i'], hits['rank']))\n"
        "np.save(sys.argv[1], hits[order])\n" % (root, os.path.join(root, "tests")))
    outs = []
    for m in ("2", mode):
        out = str(tmp_path / ("hits_mode%s.npy" % m))
        env = dict(os.environ, FR3D_CLASSIFY_MODE=m)
        subprocess.run([sys.executable, "-c", code, out], check=True, env=env, cwd=root)
        outs.append(np.load(out))
    a, b = outs
    assert len(a) > 20000 and len(a) == len(b)
    for f in ("i", "j", "rank", "slot", "code", "box"):
        assert np.array_equal(a[f], b[f]), f
    for f in ("cutoff_distance", "max_gap"):          # the warp-per-pair kernel sums the box terms in another fma pattern
        assert np.abs(a[f] - b[f]).max() < 1e-12, f


@pytest.mark.parametrize("shift", [0.0, 24000.0])
def test_single_precision_screen_is_a_superset(shift, tmp_path):
    """The fp32 screen (classify_screen32.cuh) must let through every pair the fp64 screen lets through, and the
    fp32 stacking kernel (classify_stack32.cuh) must hand every pair it cannot decide with certainty to the fp64
    kernel - so the hit records are bit-identical to the all-fp64 path (FR3D_SCREEN=64) - also when the structures sit 24 000 A
    from the origin, where a plain fp32 coordinate would carry 2e-3 A (the records hold fixed-point centres and
    centre-relative coordinates instead).  FR3D_SCREEN is read once per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from conftest import ALL9\n"
        "from fr3d_python_b200 import synthetic\n"
        "from fr3d_python_b200.classifiers import NA_pairwise_interactions as NA\n"
        "from fr3d_python_b200.engine import Engine\n"
        "batch = synthetic.make_structure_batch(synthetic.pack_base(), 120, seed0=9100)\n"
        "batch.base_xyz += %r; batch.bb_xyz += %r\n"
        "bt, _ = NA._box_table(NA._default_focused(ALL9['basepair']), NA.load_ideal_basepair_hydrogen_bonds())\n"
        "hits, n_pairs, ex = Engine.get().annotate_batch(batch, bt, NA.category_mask(ALL9))\n"
        "order = np.lexsort((hits['slot'], hits['j'], hits['i'], hits['rank']))\n"
        "np.save(sys.argv[1], hits[order])\n"
        "np.save(sys.argv[1] + '.lists.npy', ex['plan'].counters.cpu().numpy())\n"
        "plan = ex['plan']; al = lambda v: (v + 255) // 256 * 256\n"
        "off = 6 * al(4 * plan.pair_cap) + al(320 * batch.n)\n"
        "np.save(sys.argv[1] + '.unc.npy', plan.cws[off:off + 8].cpu().numpy().view(np.int64))\n"
        % (root, os.path.join(root, "tests"), shift, shift))
    outs, lists = [], []
    for m in ("32", "64"):
        out = str(tmp_path / ("hits_screen%s.npy" % m))
        env = dict(os.environ, FR3D_SCREEN=m)
        subprocess.run([sys.executable, "-c", code, out], check=True, env=env, cwd=root)
        outs.append(np.load(out))
        lists.append(np.load(out + ".lists.npy"))
        if m == "32":
            undecided = int(np.load(out + ".unc.npy")[0])
    a, b = outs
    assert len(a) > 50000 and len(a) == len(b)
    assert a.tobytes() == b.tobytes()
    assert lists[0][0] == lists[1][0] and lists[0][1] == lists[1][1]            # pairs, hits
    assert np.all(lists[0][4:8] >= lists[1][4:8])                               # every work list at least as long ...
    assert np.all(lists[0][4:8] <= 1.02 * lists[1][4:8] + 16)                   # ... and the margins cost < 2 %
    # stacking list: decided in single precision except where a comparison is within its margin (classify_stack32.cuh)
    assert 0 < undecided < 0.05 * lists[0][5]


@pytest.mark.parametrize("case", ["detail", "focus_cww_ths_stacking", "focus_cww_coplanar_covalent"])
def test_category_subsets_and_focused_interactions(golden, case):
    """Category subsets (work lists of families that were not asked for are not launched) and focused
    interaction lists (smaller cutoff-box tables) against rows produced by the unmodified reference."""
    spec = golden("spec_1gid_full.json.gz")
    gg = golden("annot_1gid_full_category_subsets.json.gz")
    g, uid = gg[case], gg["unit_ids"]
    want = {k: [(uid[a], uid[b], c) for a, b, c in v] for k, v in g["interaction_to_list_of_tuples"].items()}
    wantc = {k: set(v) for k, v in g["category_to_interactions"].items()}
    i2t, c2i = as_plain(NA.annotate_batch(spec, g["categories"])[0])
    assert i2t == want and c2i == wantc
    st = Structure.from_spec(spec)
    r = NA.annotate_nt_nt_in_structure(st, g["categories"])
    assert as_plain(r) == (want, wantc)


# ------------------------------------------------------------------ glycosidic bond orientation (N4)
@pytest.mark.parametrize("name", ["1gid_full", "1gid_modified", "1gid_variants", "1s72_strand", "chi_sweep"])
def test_bond_orientation_identical_to_reference(golden, name):
    """fr3d_bond_orientation through the reference's entry point against the rows of the unmodified reference."""
    from fr3d_python_b200.classifiers import NA_unit_annotation as UA
    from fr3d_python_b200.data import Structure
    g = golden("glycosidic.json.gz")[name]
    spec = golden("spec_%s.json.gz" % name)
    (rows, errors), = UA.annotate_bond_orientation_batch(spec, pipeline=True)
    assert rows == g["rows"]
    assert errors == g["errors"]
    # ... the same through a Structure object, and the torsion itself against the oracle's double
    rows2, errors2 = UA.annotate_bond_orientation(Structure.from_spec(spec), pipeline=True)
    assert rows2 == g["rows"] and errors2 == g["errors"]
    chi, cls = UA.bond_orientation_arrays(packing.pack_specs(spec))
    want = O.bond_orientation(spec)
    for k, w in enumerate(want):
        assert UA.ORIENTATION[cls[k]] == w["orientation"]
        if w["chi"] is not None:
            assert abs(chi[k] - w["chi"]) < 1e-9
        else:
            assert np.isnan(chi[k])


def test_bond_orientation_on_a_large_batch():
    """1000 structures in one launch: classes consistent with chi, every nt with its four atoms gets a value."""
    from fr3d_python_b200.classifiers import NA_unit_annotation as UA
    batch = synthetic.make_structure_batch(synthetic.pack_base(), 1000, seed0=4000)
    chi, cls = UA.bond_orientation_arrays(batch)
    assert len(chi) == batch.n and not np.isnan(chi).any() and np.all(cls > 0)
    want = np.where((chi > -90) & (chi < -45), 3, np.where((chi >= -45) & (chi < 90), 2, 1))
    assert np.array_equal(cls, want)
    one = UA.bond_orientation_arrays(packing.select_structures(batch, [517]))
    lo, hi = batch.struct_off[517], batch.struct_off[518]
    assert np.array_equal(one[0], chi[lo:hi]) and np.array_equal(one[1], cls[lo:hi])


# ------------------------------------------------------------------ FR3D search screening (N3)
@pytest.mark.parametrize("name", ["1gid_full", "1gid_variants"])
def test_search_screen_identical_to_reference(golden, name):
    """fr3d_radius_search through the reference's entry points against the unmodified reference's output: same
    triples in the same orientation and order; squared distances to the last bit or one (the reference squares
    with libm's pow, the device multiplies)."""
    from fr3d_python_b200.search import pair_processing as PP
    case = golden("search_screen.json.gz")[name]
    centers = np.array([[np.nan if v is None else v for v in c] for c in case["centers"]], dtype=float)
    models = case["models"]
    for cutoff, want in case["searches"].items():
        got = PP.fixed_radius_search(len(centers), models, centers, float(cutoff), 100.0)
        assert [(a, b) for a, b, _ in got] == [(a, b) for a, b, _ in want]
        g, w = np.array([c for _, _, c in got]), np.array([c for _, _, c in want])
        assert np.abs(g - w).max() <= 4e-16 * w.max()
    pl = PP.get_pairlist(case["Q"], models, centers)
    got = {"%d,%d" % (i, j): [list(p) for p in v] for i, d in pl.items() for j, v in d.items()}
    assert got == case["pairlist"]
    # the alt_index form (centres of the listed positions only, indices mapped through alt_index, :78-88) and the
    # symbolic form against the oracle
    alt = [7, 3, 5]
    few = centers[np.isfinite(centers[:, 0])][:3]
    want_alt = O.get_pairlist(case["Q"], models[:3], few, alt)
    got_alt = PP.get_pairlist(case["Q"], models[:3], few, alt)
    assert sum(len(v) for d in want_alt.values() for v in d.values()) > 0
    assert {k: dict(v) for k, v in got_alt.items()} == {k: dict(v) for k, v in want_alt.items()}
    assert PP.get_pairlist(dict(case["Q"], type="symbolic"), models, centers)[0][1] == "full"


def test_search_screen_large_and_edge_cases():
    """25 280 points (C3) against a brute-force count on a slice; empty input; all points without a centre."""
    from fr3d_python_b200.search import pair_processing as PP
    big = synthetic.make_tiled_structure(synthetic.pack_base(), 80, (5, 4, 4), seed=25000, pdb="C3")
    gframes.fit_batch_frames(big)
    a, b, dd = PP.radius_pairs(big.center, np.zeros(big.n, dtype=np.int32), 16.0)
    assert len(a) > 100000 and dd.max() < 256.0 and np.all(a != b)
    und = np.minimum(a, b).astype(np.int64) * big.n + np.maximum(a, b)
    assert len(np.unique(und)) == len(und)                          # every unordered pair once
    sub = big.center[:2000]
    d2 = ((sub[:, None] - sub[None]) ** 2).sum(axis=2)
    want = int((np.triu(d2 < 256.0, 1)).sum())
    inside = (a < 2000) & (b < 2000)
    assert int(inside.sum()) == want
    assert np.abs(dd - ((big.center[a] - big.center[b]) ** 2).sum(axis=1)).max() < 1e-10
    assert PP.fixed_radius_search(0, [], np.zeros((0, 3)), 10.0, 100.0) == []
    assert PP.fixed_radius_search(5, ["1"] * 5, np.full((5, 3), np.nan), 10.0, 100.0) == []
    two_models = PP.fixed_radius_search(4, ["1", "1", "2", "2"], np.array([[0., 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]]), 5.0, 10.0)
    assert sorted((a, b) for a, b, _ in two_models) == [(0, 1), (2, 3)]     # never across models


def test_search_screen_capacity_overflow_is_rerun_not_truncated():
    """A dense cloud yields more pairs than the first output capacity (64 n + 1024): the call is repeated with the
    exact count (FR3D_E_CAPACITY semantics), nothing is dropped."""
    from fr3d_python_b200.search import pair_processing as PP
    rng = np.random.default_rng(3)
    pts = rng.uniform(0.0, 9.0, size=(1500, 3))
    a, b, dd = PP.radius_pairs(pts, np.zeros(len(pts), dtype=np.int32), 20.0)
    assert len(a) == 1500 * 1499 // 2 > 64 * 1500 + 1024
    und = np.minimum(a, b).astype(np.int64) * 1500 + np.maximum(a, b)
    assert len(np.unique(und)) == len(und)
    assert np.abs(dd - ((pts[a] - pts[b]) ** 2).sum(axis=1)).max() < 1e-12


def test_two_devices_from_two_threads_of_one_process():
    """The header's multi-device promise (state is kept per device): two threads, one per GPU, annotate different
    batches at the same time; each result equals the single-device answer.  Needs two visible GPUs."""
    import threading
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    base = synthetic.pack_base()
    batches = [synthetic.make_structure_batch(base, 12, seed0=300 + 100 * d) for d in range(2)]
    want = [NA.annotate_batch(b, ALL9, device=0, as_arrays=True) for b in batches]
    got, errors = [None, None], []

    def work(d):
        try:
            torch.cuda.set_device(d)
            for _ in range(3):
                got[d] = NA.annotate_batch(batches[d], ALL9, device=d, as_arrays=True)
        except Exception as ex:          # surfaced in the main thread
            errors.append((d, repr(ex)))

    threads = [threading.Thread(target=work, args=(d,)) for d in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for d in range(2):
        assert got[d].n_pairs == want[d].n_pairs
        for s in range(12):
            assert np.array_equal(got[d].rows_of(s), want[d].rows_of(s)), (d, s)
    # a device that was never initialised is refused, not served with another device's tables
    from fr3d_python_b200 import _lib as L
    lib = L.load()
    if torch.cuda.device_count() > 2:
        with torch.cuda.device(2):
            assert lib.fr3d_pair_search_workspace(10) > 0                      # pure size query: fine anywhere
            rc = lib.fr3d_frames_fit(None, None, 0, None)
            assert rc == L.E_ARG


def test_cif_text_to_annotations(golden):
    """mmCIF text in, reference-format annotations out: the CIF scenario (DNA, modified residues, alternate conformers,
    two models, three symmetry operators) through cif.reader -> annotate_batch equals the oracle on the residues the
    unmodified reference reader produced from the same file."""
    import gzip
    import os
    from conftest import GOLD
    from fr3d_python_b200.cif import reader
    with gzip.open(os.path.join(GOLD, "cif_scenario.cif.gz"), "rt") as f:
        text = f.read()
    ref_spec = golden("cif_scenario_reference_reader.json.gz")
    want, wc, _, _ = oracle_tuples(ref_spec, ALL9)
    got = NA.annotate_batch(reader.cif_to_batch(text), ALL9)
    assert as_plain(got[0]) == (want, wc)
    assert sum(len(v) for v in want.values()) > 500


def test_native_cif_ingest_to_annotations(golden):
    """The C++ ingest in front of the GPU path: same annotations as the Python reader in front of it."""
    import gzip
    import os
    from conftest import GOLD
    from fr3d_python_b200.cif import native, reader
    with gzip.open(os.path.join(GOLD, "cif_scenario.cif.gz"), "rt") as f:
        text = f.read()
    a = NA.annotate_batch(native.cif_to_batch([text, text], n_threads=2), ALL9)
    b = NA.annotate_batch(reader.cif_to_batch(text), ALL9)
    assert len(a) == 2
    assert as_plain(a[0]) == as_plain(b[0]) == as_plain(a[1])


def test_large_structures_device_tail_equals_host_tail():
    """The tail's large-structure paths: S_1S72_like (2 844 nt, ~7 k hits: key-only network over the global workspace)
    and C3 (25 280 nt in ONE structure, ~63 k hits: the packed hit key no longer fits 64 bits, so the (key, value)
    network runs; 160 chains) against the host form of the tail - rows, order, dict key order."""
    import time
    base = synthetic.pack_base()
    for copies, lattice, seed in ((9, (3, 3, 1), 1072), (80, (5, 4, 4), 25000)):
        big = synthetic.make_tiled_structure(base, copies, lattice, seed=seed, pdb="BIG%d" % copies)
        t0 = time.perf_counter()
        dev = NA.annotate_batch(big, ALL9, as_arrays=True)
        t_dev = time.perf_counter() - t0
        host = NA.annotate_batch(big, ALL9, tail="host")
        d = dev[0]
        assert list(d[0].keys()) == list(host[0][0].keys())
        assert as_plain(d) == as_plain(host[0])
        assert dev.n_rows > 5 * big.n
        print("%d nt: device path %.1f ms, %d rows" % (big.n, 1e3 * t_dev, dev.n_rows))


def test_mixed_batch_of_small_and_large_structures():
    """Small structures (256-thread blocks of the tail) and large ones (>= 1 024 nts: the 1 024-thread instantiation with the
    radix sorts) in ONE batch: every structure's rows equal those of the structure annotated alone with the host tail."""
    base = synthetic.pack_base()
    small = synthetic.make_structure_batch(base, 3, seed0=7100)
    big = synthetic.make_tiled_structure(base, 9, (3, 3, 1), seed=1072, pdb="BIGM")
    mid = synthetic.make_tiled_structure(base, 4, (2, 2, 1), seed=77, pdb="MIDM")          # 1 264 nt: just above the split
    specs = [synthetic.batch_to_spec(small, 0), synthetic.batch_to_spec(big, 0), synthetic.batch_to_spec(small, 1),
             synthetic.batch_to_spec(mid, 0), synthetic.batch_to_spec(small, 2)]
    got = NA.annotate_batch(specs, ALL9, as_arrays=True)
    assert len(got) == 5
    for s, spec in enumerate(specs):
        alone = NA.annotate_batch([spec], ALL9, tail="host")
        assert list(got[s][0].keys()) == list(alone[0][0].keys()), s
        assert as_plain(got[s]) == as_plain(alone[0]), s


def test_large_structures_in_ordinary_tail_blocks(monkeypatch):
    """Many large structures in one call are not announced (engine.DeviceIdent, FR3D_TAIL_BIG_LIMIT) and run in the
    256-thread blocks, whose sorts beyond 2 048 keys are the radix sort: forced here for the 2 844-nt and the 25 280-nt
    structure (key-only and (key, value) form), rows equal to the host tail."""
    monkeypatch.setenv("FR3D_TAIL_BIG_LIMIT", "0")
    base = synthetic.pack_base()
    for copies, lattice, seed in ((9, (3, 3, 1), 1072), (80, (5, 4, 4), 25000)):
        big = synthetic.make_tiled_structure(base, copies, lattice, seed=seed, pdb="BIGS%d" % copies)
        dev = NA.annotate_batch(big, ALL9, as_arrays=True)
        host = NA.annotate_batch(big, ALL9, tail="host")
        assert list(dev[0][0].keys()) == list(host[0][0].keys())
        assert as_plain(dev[0]) == as_plain(host[0])


def test_packed_upload_equals_plain_upload(golden):
    """Pipeline(packed_upload=True): ragged observed base atoms + backbone slots 0..8 over PCIe, the records completed on the
    device (fr3d_frames_fit_ragged, fr3d_backbone_expand) - rows identical to the plain upload, on a perturbed batch with
    missing atoms and on the variants structure (two models, symmetry copy, alternate ids: previous-O3' links that are not
    the previous row)."""
    import torch
    from fr3d_python_b200.engine import Pipeline, pin_batch
    eng = Engine.get()
    bt, atoms, labels = NA._tables(NA._default_focused(ALL9['basepair']), NA.load_ideal_basepair_hydrogen_bonds())
    mask = NA.category_mask(ALL9)
    base = synthetic.pack_base()
    batch = synthetic.make_structure_batch(base, 21, seed0=3100)
    rng = np.random.default_rng(9)
    lost = rng.random(batch.n) < 0.15
    batch.base_mask[lost] &= ~np.uint32(1 << 4)                       # a missing base atom here and there
    nobb = rng.random(batch.n) < 0.1
    batch.bb_mask[nobb] &= ~np.uint32(1 << 5)                         # ... and a missing O3' (breaks the link of the next row)
    variants = packing.pack_specs(golden("spec_1gid_variants.json.gz"))
    for b in (batch, variants):
        pinned = pin_batch(torch, b)
        if pinned["base_xyz"].shape[1] >= 16:
            continue
        outs = []
        for packed in (False, True):
            pipe = Pipeline(eng, b, pinned, bt, mask, n_chunks=3, labels=labels, packed_upload=packed)
            assert (pipe.chunks[0].packed is not None) == packed
            for _ in range(2):
                out = pipe.run_graph()
            outs.append((out.n_pairs, out.rows.copy(), out.row_start.copy(), out.row_count.copy(), pipe.h2d_bytes))
        assert outs[0][0] == outs[1][0]
        for s in range(b.n_structures):
            r0 = outs[0][1][outs[0][2][s]:outs[0][2][s] + outs[0][3][s]]
            r1 = outs[1][1][outs[1][2][s]:outs[1][2][s] + outs[1][3][s]]
            assert np.array_equal(r0, r1), s
        assert outs[1][4] < 0.93 * outs[0][4]
