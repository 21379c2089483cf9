"""Host side of the product (packing, cutoff flattening, post-processing, writers, synthetic
inputs) on CPU.  The GPU stage is replaced by hit records made from the oracle's per-rule
traces, so the order-dependent tail is checked against the reference golden without a GPU."""

import copy
import os

import numpy as np
import pytest

from conftest import ALL9
from fr3d_python_b200 import _lib, packing, postprocess, synthetic
from fr3d_python_b200 import definitions as defs
from fr3d_python_b200.classifiers import NA_pairwise_interactions as NA
from fr3d_python_b200.engine import BoxTable, HIT_DTYPE, reverse_edges
from oracle import fr3d_oracle as O

_SO = {n: k for k, n in enumerate(["O2'", "O3'", "O4'", "O5'", "OP1", "OP2"])}
_ST = {"s35": 0, "s53": 1, "s33": 2, "s55": 3}


def traces_to_hits(spec, tr, bt):
    """Encode the oracle's per-rule results as the GPU would (fr3d_hit_t), in scrambled order."""
    nts = O.build_nts(spec)
    cubes, _ = O.make_nt_cubes_half(nts)
    first = {}
    cell_of = {}
    for key, members in cubes.items():
        for m in members:
            cell_of[m] = key
        first[key] = min(members)
    rank_of = {}
    for i, j in tr["candidates"]:
        k1, k2 = cell_of[i], cell_of[j]
        off = [k2[0] - k1[0], k2[1] - k1[1], k2[2] - k1[2]]
        rank_of[(i, j)] = first[k1] * 16 + O.HALF_STENCIL.index(off)
    rows = []

    def add(i, j, slot, code, box=0, cd=0.0, gap=0.0):
        rows.append((i, j, rank_of[(i, j)], slot, code, box, cd, gap))

    cand = set(map(tuple, tr["candidates"]))
    for a, b, lab in tr["sO"]:
        if lab:
            near = lab.startswith("n")
            body = lab[2:] if near else lab[1:]
            code = _SO[body[1:]] | (8 if body[0] == "5" else 0) | (16 if near else 0)
            if (a, b) in cand:
                add(a, b, _lib.SLOT_SO12, code)
            else:
                add(b, a, _lib.SLOT_SO21, code)
    for a, b, lab in tr["stacking"]:
        if lab:
            add(a, b, _lib.SLOT_STACK, _ST[lab.replace("n", "")] | (4 if lab.startswith("n") else 0))
    for a, b, lab in tr["sugar_ribose"]:
        if lab:
            code = 0 if lab == "cSR" else 1
            if (a, b) in cand:
                add(a, b, _lib.SLOT_SR12, code)
            else:
                add(b, a, _lib.SLOT_SR21, code)
    for a, b, bph, br in tr["backbone"]:
        if bph:
            add(a, b, _lib.SLOT_BPH, int(bph.replace("n", "")[0]) | (16 if bph.startswith("n") else 0))
        if br:
            add(a, b, _lib.SLOT_BR, int(br.replace("n", "")[0]) | (16 if br.startswith("n") else 0))
    cp_done = set()
    for a, b, cp, gap12, mind, val in tr["coplanar"]:
        pair = (a, b) if (a, b) in cand else (b, a)
        if cp and pair not in cp_done:
            cp_done.add(pair)
            add(pair[0], pair[1], _lib.SLOT_CP, 0)
    # basepairs: replay the 12-vs-21 arbitration (NA:935-957) on the oracle's per-orientation results
    per = {}
    for a, b, lw, sub, cd, gap in tr["basepair"]:
        pair, ori = ((a, b), 0) if (a, b) in cand else ((b, a), 1)
        per.setdefault(pair, {})[ori] = (lw, sub, cd, gap)
    for pair, d in per.items():
        r12 = d.get(0, ("", None, None, None))
        r21 = d.get(1, ("", None, None, None))
        i12, i21 = r12[0], r21[0]
        if i12 and i21:
            if reverse_edges(i12).lower() == i21.lower():
                i21 = ""
            elif "n" in i12 and "n" in i21:
                if r12[2] < r21[2]:
                    i21 = ""
                else:
                    i12 = ""
            elif "n" in i21:
                i21 = ""
            elif "n" in i12:
                i12 = ""
            else:
                i21 = ""
        if i12:
            use, r = 0, r12
        elif i21:
            use, r = 1, r21
        else:
            continue
        nt_a = nts[pair[0] if use == 0 else pair[1]]
        nt_b = nts[pair[1] if use == 0 else pair[0]]
        combo = O.get_parent(nt_a.sequence) + "," + O.get_parent(nt_b.sequence)
        name = r[0][1:] if r[0].startswith("n") else r[0]
        box = [k for k in range(len(bt.box_name)) if bt.box_combo[k] == combo and bt.box_name[k] == name
               and bt.box_subcat[k] == r[1]]
        assert box, (combo, name, r)
        add(pair[0], pair[1], _lib.SLOT_BP, (1 if r[0].startswith("n") else 0) | (use << 1), box[0], r[2], r[3])
    hits = np.array(rows, dtype=HIT_DTYPE)
    rng = np.random.default_rng(0)
    return hits[rng.permutation(len(hits))]


@pytest.mark.parametrize("spec_name,annot_name", [
    ("spec_1s72_strand.json.gz", "annot_1s72_strand_all9.json.gz"),
    ("spec_1gid_full.json.gz", "annot_1gid_full_all9.json.gz"),
])
def test_postprocessing_reproduces_reference_output(golden, spec_name, annot_name):
    spec = golden(spec_name)
    g = golden(annot_name)
    cats = g["categories"]
    _, _, tr = O.annotate(spec, cats)
    focused = NA.focus_basepair_cutoffs(NA.nt_nt_cutoffs, cats["basepair"])
    bt, atoms = NA._box_table(focused, NA.load_ideal_basepair_hydrogen_bonds())
    hits = postprocess.sort_hits(traces_to_hits(spec, tr, bt))
    batch = packing.pack_specs(spec)
    out, c2i = postprocess.build_results(batch, 0, hits, bt, atoms, cats)
    uid = g["unit_ids"]
    want = {k: [(uid[a], uid[b], c) for a, b, c in v] for k, v in g["interaction_to_list_of_tuples"].items()}
    assert {k: v for k, v in out.items()} == want              # identical rows in identical order
    assert {k: sorted(v) for k, v in c2i.items() if v} == g["category_to_interactions"]


def test_packing_layout(golden):
    spec = golden("spec_1gid_full.json.gz")
    b = packing.pack_specs(spec)
    assert b.n == 316 and b.n_structures == 1
    assert b.unit_ids()[0] == "1GID|1|A|G|103"
    g = b.meta[0]
    assert g[0] == defs.PARENT_CODE['G'] and g[1] & _lib.FLAG_STD_RNA
    assert b.base_mask[0] & 0xFFFF == (1 << 11) - 1           # 11 heavy atoms of G, no hydrogens yet
    assert b.bb_mask[0] == (1 << 9) - 1              # first nt of the chain: no previous O3'
    assert b.bb_mask[1] == (1 << 10) - 1
    assert np.array_equal(b.bb_xyz[1, defs.BB_PREV_O3], b.bb_xyz[0, 5])
    # glycosidic atom is slot 0 for every parent
    for k in range(5):
        nt = spec["nts"][k]
        want = "N9" if nt["sequence"] in "AG" else "N1"
        xyz = [a[1:] for a in nt["atoms"] if a[0] == want][0]
        assert np.array_equal(b.base_xyz[k, 0], xyz)
    # chain B starts a new chain: no previous O3'
    first_b = b.chain.index("B")
    assert not (b.bb_mask[first_b] >> defs.BB_PREV_O3 & 1)


def test_packing_missing_and_duplicate_atoms():
    nt = {"chain": "A", "sequence": "U", "number": "5", "index": 5, "model": "1", "symmetry": "1_555",
          "alt_id": None, "insertion_code": None,
          "atoms": [["N1", 0, 0, 0], ["C2", 1, 0, 0], ["C2", 3, 0, 0], ["XX", 9, 9, 9], ["O2'", 1, 1, 1]]}
    unknown = dict(nt, sequence="HOH", number="6", index=6)
    b = packing.pack_specs({"pdb": "T", "nts": [nt, unknown]})
    assert b.base_mask[0] & 0xFFFF == 0b11 and np.array_equal(b.base_xyz[0, 1], [2, 0, 0])    # duplicates averaged
    assert b.bb_mask[0] == 1 << 6
    assert b.meta[1, 0] == -1 and b.base_mask[1] == 0


def test_unit_id_encoding():
    f = packing.encode_unit_id
    assert f("1GID", "1", "A", "G", "103", None, None, "1_555") == "1GID|1|A|G|103"
    assert f("1GID", 1, "A", "G", 103, None, None, "6_555") == "1GID|1|A|G|103||||6_555"
    assert f("2XYZ", "1", "A", "C", "7", "B", None, "1_555") == "2XYZ|1|A|C|7||B"
    assert f("2XYZ", "1", "A", "C", "7", None, "A", "1_555") == "2XYZ|1|A|C|7|||A"
    nts = O.build_nts({"pdb": "2XYZ", "nts": [{"chain": "A", "sequence": "C", "number": "7", "index": 7,
                                             "model": "1", "symmetry": "2_655", "alt_id": "B",
                                             "insertion_code": None, "atoms": []}]})
    assert nts[0].unit_id() == f("2XYZ", "1", "A", "C", "7", "B", None, "2_655")


def test_box_table_matches_focus():
    focused = NA.focus_basepair_cutoffs(NA.nt_nt_cutoffs, NA.Leontis_Westhof_basepairs)
    ofocused = O.focus_basepair_cutoffs(O.nt_nt_cutoffs(), NA.Leontis_Westhof_basepairs)
    assert focused == ofocused
    bt = BoxTable(focused)
    for a, pa in enumerate(defs.PARENTS):
        for b, pb in enumerate(defs.PARENTS):
            combo = pa + "," + pb
            for sg, key in enumerate((1, -1)):
                start, count = bt.index[a * 5 + b, sg]
                if combo not in defs.BASEPAIR_COMBOS:
                    assert count == 0
                    continue
                want = [(i, s) for i, subs in focused[combo][key].items() for s in subs]
                got = list(zip(bt.box_name[start:start + count], bt.box_subcat[start:start + count]))
                assert got == want
                for k in range(start, start + count):
                    cut = focused[combo][key][bt.box_name[k]][bt.box_subcat[k]]
                    assert bt.boxes['lim'][k][0] == cut['xmin'] and bt.boxes['lim'][k][10] == cut['gapmax']
                    assert bool(bt.boxes['flags'][k] & 8) == ('radiusmax' in cut)
    # near ("n...") and alternative ("a...") tables are never selected
    assert not any(n[0] in "na" for n in bt.box_name)


def test_reverse_edges():
    for a, b in [("cWW", "cWW"), ("tHS", "tSH"), ("ntHS", "ntSH"), ("cWSa", "cSWa"), ("ncWSa", "ncSWa"),
                 ("cp", "cp"), ("s35", "s53"), ("cSR", "cRS"), ("ns35", "ns53")]:
        assert reverse_edges(a) == b and O.reverse_edges(a) == b


def test_same_edge_cleanup_cases():
    hb = frozenset(["H1", "N1"])
    # two true pairs on the same edge sharing a hydrogen: the larger max_gap is demoted
    u2bp = {0: [["cWW", (0.0, 0.2, hb), 1], ["cWH", (0.0, 0.9, hb), 2]],
            1: [["cWW", (0.0, 0.2, hb), 0]], 2: [["cHW", (0.0, 0.9, frozenset(["N7"])), 0]]}
    rem, near = postprocess.same_edge_cleanup(u2bp)
    assert rem == set() and near == {(0, 2), (2, 0)}
    # true + bad near: near removed
    u2bp = {0: [["cWW", (0.0, 0.2, hb), 1], ["ncWH", (0.8, 0.9, hb), 2]], 1: [], 2: []}
    rem, near = postprocess.same_edge_cleanup(u2bp)
    assert rem == {(0, 2), (2, 0)} and near == set()
    # different edges: untouched
    u2bp = {0: [["cWW", (0.0, 0.2, hb), 1], ["cHW", (0.0, 0.9, hb), 2]]}
    assert postprocess.same_edge_cleanup(u2bp) == (set(), set())


def test_writers_and_loader(tmp_path, golden):
    g = golden("annot_1s72_strand_all9.json.gz")
    uid = g["unit_ids"]
    i2t = {k: [(uid[a], uid[b], c) for a, b, c in v] for k, v in g["interaction_to_list_of_tuples"].items()}
    c2i = {k: set(v) for k, v in g["category_to_interactions"].items()}
    from collections import defaultdict
    c2i = defaultdict(set, c2i)
    cats = {"stacking": [], "sO": [], "basepair": NA.Leontis_Westhof_basepairs}
    NA.write_txt_output_file(str(tmp_path), "1S72", i2t, cats, c2i)
    rows = open(os.path.join(str(tmp_path), "1S72_stacking.txt")).read().splitlines()
    assert len(rows) == 16 and rows[0].split("\t") == ["1S72|1|0|A|11", "s35", "1S72|1|0|U|12", "0"]
    nums = [int(r.split("\t")[0].split("|")[4]) for r in rows]
    assert nums == sorted(nums)
    assert open(os.path.join(str(tmp_path), "1S72_sO.txt")).read().count("\n") == 2


def test_synthetic_generators_are_deterministic():
    base = synthetic.pack_base()
    a = synthetic.make_structure_batch(base, 3, seed0=4000)
    b = synthetic.make_structure_batch(base, 3, seed0=4000)
    assert np.array_equal(a.base_xyz, b.base_xyz) and np.array_equal(a.bb_xyz, b.bb_xyz)
    one = synthetic.make_structure_batch(base, 1, seed0=4002)
    assert np.array_equal(one.base_xyz, a.base_xyz[632:])           # structure k depends only on seed0 + k
    assert a.n_structures == 3 and a.unit_ids()[316] == "S0001|1|A|G|103"
    spec = synthetic.batch_to_spec(a, 1)
    again = packing.pack_specs(spec)
    assert np.array_equal(again.base_xyz, a.base_xyz[316:632]) and np.array_equal(again.bb_mask, a.bb_mask[316:632])
    assert np.array_equal(again.bb_xyz, a.bb_xyz[316:632])          # incl. the previous-O3' slot
    t = synthetic.make_tiled_structure(base, 4, (2, 2, 1), 7)
    assert t.n == 4 * 316 and len(set(t.chain)) == 8 and t.n_structures == 1
    spec = synthetic.batch_to_spec(t, 0)
    again = packing.pack_specs(spec)
    assert np.array_equal(again.meta[:, 3], t.meta[:, 3]) and np.array_equal(again.bb_xyz, t.bb_xyz)


def test_select_structures_roundtrip():
    base = synthetic.pack_base()
    a = synthetic.make_structure_batch(base, 4, seed0=10)
    s = packing.select_structures(a, [2, 0])
    assert s.pdb == ["S0002", "S0000"] and s.n == 632
    assert np.array_equal(s.base_xyz[:316], a.base_xyz[632:948])
    assert s.unit_ids()[316] == a.unit_ids()[0]


def test_no_cpu_fallback():
    """The product path must fail loudly without a CUDA device."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.Fr3dError):
        NA.annotate_batch(synthetic.load_golden_spec("spec_1s72_strand"), ALL9)
    from fr3d_python_b200.geometry import discrepancy as disc
    with pytest.raises(_lib.Fr3dError):
        disc.matrix_discrepancy([np.zeros(3)] * 3, [np.eye(3)] * 3, [np.zeros(3)] * 3, [np.eye(3)] * 3)


def test_product_does_not_import_oracle():
    """Only tests/, smoke() and bench.py's CPU legs may touch oracle/."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "fr3d_python_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle|fr3d_oracle|import_module\(.oracle", re.M)
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(d, f)).read()
                assert not pat.search(text), os.path.join(d, f)


def test_host_tail_in_forked_workers_equals_serial(golden):
    """NA._host_tail (the per-structure tail of annotate_batch) with worker processes: 24 copies of the 1S72 strand
    as one batch, hit records encoded from the oracle's traces with global nt indices, like the GPU's."""
    spec = golden("spec_1s72_strand.json.gz")
    cats = golden("annot_1s72_strand_all9.json.gz")["categories"]
    _, _, tr = O.annotate(spec, cats)
    focused = NA.focus_basepair_cutoffs(NA.nt_nt_cutoffs, cats["basepair"])
    bt, atoms = NA._box_table(focused, NA.load_ideal_basepair_hydrogen_bonds())
    one = traces_to_hits(spec, tr, bt)
    S, n = 24, len(spec["nts"])
    specs = [dict(spec, pdb="X%03d" % k) for k in range(S)]
    batch = packing.pack_specs(specs)
    parts = []
    for k in range(S):
        h = one.copy()
        h["i"] += k * n
        h["j"] += k * n
        h["rank"] += 16 * k * n
        parts.append(h)
    hits = postprocess.sort_hits(np.concatenate(parts))
    sidx = np.searchsorted(batch.struct_off, hits["i"], side="right") - 1
    cuts = np.searchsorted(sidx, np.arange(S + 1))
    serial = NA._host_tail(batch, hits, cuts, bt, atoms, cats, 1)
    forked = NA._host_tail(batch, hits, cuts, bt, atoms, cats, 3)
    assert len(serial) == len(forked) == S
    for k, (a, b) in enumerate(zip(serial, forked)):
        assert list(a[0].keys()) == list(b[0].keys())
        assert {x: list(v) for x, v in a[0].items()} == {x: list(v) for x, v in b[0].items()}
        assert {x: set(v) for x, v in a[1].items()} == {x: set(v) for x, v in b[1].items()}
        assert all(u.startswith("X%03d|" % k) for v in a[0].values() for u, _, _ in v)
    assert NA._TAIL_JOB is None


@pytest.mark.parametrize("name", ["1gid_full", "1gid_modified", "1gid_variants", "1s72_strand", "chi_sweep"])
def test_array_native_packer_equals_pack_specs(golden, name):
    """packing.pack_columns (no Python loop over atoms) against packing.pack_specs on every fixture: DNA, modified
    residues with and without observed hydrogens, missing atoms, alt ids, two models, symmetry copies."""
    spec = golden("spec_%s.json.gz" % name)
    if name == "1s72_strand":                # observed (mislabelled) amino hydrogens and a repeated atom name
        spec = copy.deepcopy(spec)
        for nt in spec["nts"]:
            xyz = {a[0]: a[1:] for a in nt["atoms"]}
            if nt["sequence"] == "A":
                nt["atoms"] += [["H61"] + [v + 0.9 for v in xyz["N6"]], ["H62"] + [v - 0.9 for v in xyz["N6"]]]
            if nt["sequence"] == "U":
                nt["atoms"] += [["O4"] + [v + 0.2 for v in xyz["O4"]]]
    want = packing.pack_specs(spec)
    got = packing.pack_columns(*packing.specs_to_columns(spec))
    for field in ("center", "rot", "base_xyz", "bb_xyz", "base_mask", "bb_mask", "meta", "struct_off"):
        assert np.array_equal(getattr(got, field), getattr(want, field)), field
    for field in ("pdb", "model", "chain", "sequence", "number", "index", "symmetry", "alt_id", "insertion_code"):
        assert getattr(got, field) == getattr(want, field), field
    assert got.unit_ids() == want.unit_ids() and got.frames_given == want.frames_given


def test_array_native_packer_on_a_batch_and_empty_input():
    base = synthetic.pack_base()
    batch = synthetic.make_structure_batch(base, 7, seed0=11)
    specs = [synthetic.batch_to_spec(batch, s) for s in range(7)]
    want = packing.pack_specs(specs)
    got = packing.pack_columns(*packing.specs_to_columns(specs))
    for field in ("base_xyz", "bb_xyz", "base_mask", "bb_mask", "meta", "struct_off"):
        assert np.array_equal(getattr(got, field), getattr(want, field)), field
    empty = packing.pack_columns(*packing.specs_to_columns([]))
    assert empty.n == 0 and empty.n_structures == 0


# ---- the packer's identity bookkeeping: the first (loop per nt, sort per structure) implementation as the expected
# ---- behaviour of the vectorised one
def _finish_identity_v0(batch):
    """group ids, chain ranks, alt-id keys and previous-O3' (map_unit_id_to_previous_O3, NA:480-525)."""
    n = batch.n
    sof = batch.structure_of() if n else np.zeros(0, dtype=np.int64)
    groups = {}
    chains = sorted(set(batch.chain))
    chain_rank = {c: r for r, c in enumerate(chains)}
    any_alt = any(a is not None for a in batch.alt_id)
    resid = {}
    alts = {}
    # plain Python lists inside the loops, one numpy write per column at the end
    sof_l = sof.tolist()
    m2, m3, m4, m5, m6 = [0] * n, [0] * n, [0] * n, [0] * n, [0] * n
    for k in range(n):
        m2[k] = groups.setdefault((sof_l[k], batch.model[k]), len(groups))
        m3[k] = chain_rank[batch.chain[k]]
        idx = batch.index[k]
        m4[k] = 0 if idx is None else int(idx)
        if any_alt:     # alt-id twins screen, NA:706-713
            key = (sof_l[k], batch.number[k], batch.sequence[k], batch.chain[k], batch.symmetry[k],
                   batch.insertion_code[k])
            m5[k] = resid.setdefault(key, len(resid))
            m6[k] = alts.setdefault(batch.alt_id[k], len(alts))
        else:
            m5[k] = k
    if n:
        batch.meta[:, 2] = m2; batch.meta[:, 3] = m3; batch.meta[:, 4] = m4; batch.meta[:, 5] = m5; batch.meta[:, 6] = m6
    # previous O3'
    uid = batch.unit_ids()
    has_o3 = ((batch.bb_mask >> 5) & 1).tolist()
    model, symmetry, chain = batch.model, batch.symmetry, batch.chain
    src, dst = [], []
    for s in range(batch.n_structures):
        lo, hi = int(batch.struct_off[s]), int(batch.struct_off[s + 1])
        order = sorted(range(lo, hi), key=lambda k: (model[k], symmetry[k] or "", chain[k], m4[k], uid[k]))
        prev = None
        for k in order:
            if prev is not None and model[k] == model[prev] and symmetry[k] == symmetry[prev] \
                    and chain[k] == chain[prev] and m4[k] == m4[prev] + 1:
                if has_o3[prev]:                     # O3' of the previous nt
                    src.append(prev); dst.append(k)
            prev = k
    if dst:
        batch.bb_xyz[dst, defs.BB_PREV_O3] = batch.bb_xyz[src, 5]
        batch.bb_mask[dst] |= np.uint32(1 << defs.BB_PREV_O3)




def _identity_fields(batch):
    return batch.meta.copy(), batch.bb_xyz.copy(), batch.bb_mask.copy()


@pytest.mark.parametrize("name", ["1gid_full", "1gid_modified", "1gid_variants", "1s72_strand", "chi_sweep", "shuffled"])
def test_identity_bookkeeping_equals_first_implementation(golden, name):
    from fr3d_python_b200 import definitions as defs  # noqa: F401
    if name == "shuffled":            # several structures, nts out of order, repeated positions, missing O3', no index
        spec = copy.deepcopy(golden("spec_1gid_variants.json.gz"))
        rng = np.random.default_rng(5)
        specs = []
        for k in range(4):
            nts = [copy.deepcopy(spec["nts"][i]) for i in rng.permutation(len(spec["nts"]))[:150 + 40 * k]]
            for q, nt in enumerate(nts):
                if q % 11 == 0:
                    nt["atoms"] = [a for a in nt["atoms"] if a[0] != "O3'"]
                if q % 29 == 0:
                    nt["index"] = None
            specs.append({"pdb": "SH%d" % k, "nts": nts})
        specs.append({"pdb": "EMPTY", "nts": []})
    else:
        specs = [golden("spec_%s.json.gz" % name)]
    b = packing.pack_specs(specs)
    got = _identity_fields(b)
    # undo the previous-O3' slot, then run the first implementation on the same batch
    b.bb_xyz[:, defs.BB_PREV_O3] = 0.0
    b.bb_mask &= ~np.uint32(1 << defs.BB_PREV_O3)
    b.meta[:, 2:7] = -7
    _finish_identity_v0(b)
    want = _identity_fields(b)
    for g, w, what in zip(got, want, ("meta", "bb_xyz", "bb_mask")):
        assert np.array_equal(g, w), what
    assert (got[2] >> defs.BB_PREV_O3 & 1).sum() > 0 or name == "chi_sweep"


def test_packed_upload_layout_round_trips():
    """engine._pack_upload: ragged observed base atoms + group offsets, backbone slots 0..8 + previous-O3' links; unpacking
    on the host gives back every observed coordinate (what fr3d_frames_fit_ragged / fr3d_backbone_expand do on the device)."""
    from fr3d_python_b200 import engine, synthetic
    base = synthetic.pack_base()
    batch = synthetic.make_structure_batch(base, 3, seed0=11)
    rng = np.random.default_rng(5)
    b = batch.base_xyz[:, :11].copy()
    m = batch.base_mask.copy() & 0x7FF
    drop = rng.random(len(m)) < 0.2
    m[drop] &= ~np.uint32(1 << 3)                        # some residues lack an atom
    bb, bbm = batch.bb_xyz.copy(), batch.bb_mask.copy()
    odd = np.nonzero((bbm >> 9) & 1)[0][::7]
    bb[odd, 9] += 0.25                                   # previous-O3' entries that are NOT the previous row's O3'
    p = engine._pack_upload(b, m, bb, bbm)
    n = len(m)
    assert p["group_off"].shape[0] == (n + 31) // 32 + 1 and p["group_off"][-1] == p["ragged"].shape[0]
    k = 0
    for i in range(n):
        if i % 32 == 0:
            assert p["group_off"][i // 32] == k
        for s in range(11):
            if (m[i] >> s) & 1:
                assert np.array_equal(p["ragged"][k], b[i, s])
                k += 1
    assert k == p["ragged"].shape[0]
    out = np.zeros_like(bb)
    out[:, :9] = p["bb9"]
    rel = p["prev_rel"] == 1
    out[1:][rel[1:], 9] = p["bb9"][:-1][rel[1:], 5]
    out[p["extra_nt"], 9] = p["extra_xyz"]
    has9 = ((bbm >> 9) & 1).astype(bool)
    assert np.array_equal(out[has9, 9], bb[has9, 9]) and np.array_equal(out[:, :9], bb[:, :9])
    assert set(odd.tolist()) <= set(p["extra_nt"].tolist()) and len(p["extra_nt"]) < 0.2 * has9.sum()


def test_pipeline_chunk_bounds():
    """engine.chunk_bounds: equal chunks, tapered chunks (bench.py --taper), degenerate cases."""
    from fr3d_python_b200.engine import chunk_bounds
    assert chunk_bounds(1000, 8) == [0, 125, 250, 375, 500, 625, 750, 875, 1000]
    w = [0.3, 0.7, 1.2, 1.3, 1.3, 1.3, 1.3, 1.2, 0.9, 0.5]
    b = chunk_bounds(1000, 10, w)
    assert b[0] == 0 and b[-1] == 1000 and all(y > x for x, y in zip(b, b[1:]))
    sizes = np.diff(b)
    assert sizes[0] == 30 and sizes[-1] == 50 and sizes.max() == 130
    assert chunk_bounds(3, 10, w) == [0, 1, 2, 3]                       # fewer structures than chunks: the weights are dropped
    b = chunk_bounds(5, 5, [0, 0, 0, 0, 1])                             # every chunk keeps a structure
    assert b == [0, 1, 2, 3, 4, 5]
    assert chunk_bounds(7, 1, [1.0]) == [0, 7]
