"""Generate tests/golden/* by running the UNMODIFIED reference (imported from
/root/reference) on structure specs built from the reference's own test data.

DEV-CONTAINER ONLY.  Run:  python tools/make_golden.py
The committed outputs are what the CPU/GPU test-suites compare against; the
GPU box never sees /root/reference.

Outputs (all JSON, gz where large):
  spec_*.json.gz            input structures (plain coordinates + ids)
  frames_<spec>.json.gz     per-nt centre, rotation, placed hydrogens (reference Component)
  annot_<spec>.json.gz      candidate pairs in reference order, per-rule stage traces,
                            final interaction_to_list_of_tuples (ordered) per category set
  geometry.json.gz          Kabsch / discrepancy known answers + reference outputs on seeded inputs
  matlab_1gid_frames.json.gz  the Matlab golden centres/rotations (second witness)
"""

import contextlib
import gzip
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_fixtures as rf  # noqa: E402

rf.add_reference_to_path()
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

LW18 = ['cWW', 'cSS', 'cHH', 'cHS', 'cHW', 'cSH', 'cSW', 'cWH', 'cWS', 'tSS', 'tHH', 'tHS', 'tHW',
        'tSH', 'tSW', 'tWH', 'tWS', 'tWW']
BACKBONE = ["P", "OP1", "OP2", "O5'", "C5'", "C4'", "O4'", "C3'", "O3'", "C2'", "O2'", "C1'"]


def dump(name, obj):
    path = os.path.join(GOLD, name)
    data = json.dumps(obj, separators=(",", ":")).encode()
    if name.endswith(".gz"):
        with gzip.GzipFile(path, "wb", mtime=0) as f:
            f.write(data)
    else:
        with open(path, "wb") as f:
            f.write(data)
    print("wrote %s (%d bytes)" % (name, os.path.getsize(path)))


def residues(structure):
    return list(structure.residues(type=["RNA linking", "DNA linking"]))


def make_spec_1gid_full(base_spec, strand_spec):
    """S_1GID_full: template backbone transfer."""
    sb = residues(rf.ref_structure_from_spec(base_spec))
    st = residues(rf.ref_structure_from_spec(strand_spec))
    by_num = {c.number: c for c in st}
    templates = {'A': by_num['11'], 'U': by_num['12'], 'G': by_num['13'], 'C': by_num['14']}
    local = {}
    for seq, t in templates.items():
        U = np.asarray(t.rotation_matrix)
        c = np.asarray(t.centers["base"])
        local[seq] = [(n, (np.asarray(t.centers[n]) - c) @ U) for n in BACKBONE]
    nts = []
    for nt_spec, comp in zip(base_spec["nts"], sb):
        U = np.asarray(comp.rotation_matrix)
        c = np.asarray(comp.centers["base"])
        atoms = []
        for n, x in local[nt_spec["sequence"]]:
            p = c + U @ x
            atoms.append([n, float(p[0]), float(p[1]), float(p[2])])
        new = dict(nt_spec)
        new["atoms"] = atoms + nt_spec["atoms"]
        nts.append(new)
    return {"pdb": "1GID", "nts": nts}


def make_spec_modified(full_spec):
    """1GID-full with a deterministic subset of residues turned into modified nucleotides of the
    reference's atom_mappings.txt (PSU, 5MC, OMG, 2MG, 1MA, MA6, H2U ...): atom names are renamed
    through parent_atom_to_modified, extra methyl carbons are added at idealised positions, and a few
    residues carry an observed hydrogen (which switches off the reference's ideal-atom insertion,
    components.py:465-513)."""
    from fr3d.data import mapping as mp
    from fr3d import definitions as defs
    comps = residues(rf.ref_structure_from_spec(full_spec))
    nts = []

    def unit(v):
        return v / np.linalg.norm(v)

    for k, (nt, comp) in enumerate(zip(full_spec["nts"], comps)):
        seq = nt["sequence"]
        U = np.asarray(comp.rotation_matrix)
        c = np.asarray(comp.centers["base"])
        std = {n: np.array(v) for n, v in defs.NAbasecoordinates[seq].items()}
        place = lambda x: [float(v) for v in (c + U @ x)]     # noqa: E731
        new_seq, extra = None, []
        if seq == 'U' and k % 5 == 0:
            new_seq = 'PSU'
        elif seq == 'U' and k % 11 == 7:
            new_seq = 'H2U'
            extra = [['HN3'] + place(std['H3'] + np.array([0.03, -0.02, 0.01]))]       # observed hydrogen
        elif seq == 'C' and k % 4 == 1:
            new_seq = '5MC'
            extra = [['CM5'] + place(std['C5'] + 1.5 * unit(std['H5'] - std['C5']))]
            if k % 8 == 5:
                extra.append(['H6'] + place(std['H6'] + np.array([0.02, 0.02, 0.0])))   # observed hydrogen
        elif seq == 'G' and k % 6 == 2:
            new_seq = 'OMG'
            o2 = [a for a in nt["atoms"] if a[0] == "O2'"][0]
            extra = [['CM2', o2[1] + 0.9, o2[2] + 0.7, o2[3] + 0.6]]
        elif seq == 'G' and k % 6 == 4:
            new_seq = '2MG'
            extra = [['CM2'] + place(std['N2'] + 1.45 * unit(std['H21'] - std['N2']))]
        elif seq == 'A' and k % 5 == 3:
            new_seq = '1MA'
            extra = [['CM1'] + place(std['N1'] + 1.47 * unit(std['N1'] - std['C5']))]
        elif seq == 'A' and k % 7 == 5:
            new_seq = 'MA6'
            extra = [['C9'] + place(std['N6'] + 1.45 * unit(std['H61'] - std['N6'])),
                     ['C10'] + place(std['N6'] + 1.45 * unit(std['H62'] - std['N6']))]
        new = dict(nt)
        if new_seq is not None:
            p2m = mp.parent_atom_to_modified[new_seq]
            new["sequence"] = new_seq
            new["atoms"] = [[p2m.get(a[0], a[0])] + a[1:] for a in nt["atoms"]] + extra
        nts.append(new)
    return {"pdb": "1GID", "nts": nts}


def frames_golden_modified(spec):
    """Like frames_golden plus, per residue, centers[name] of every base atom name the annotator looks
    up (get_base_atom_names, NA:1574-1599) and the base points its atoms() loops see (NA:2211-2220)."""
    from fr3d.classifiers import NA_pairwise_interactions as NA
    comps = residues(rf.ref_structure_from_spec(spec))
    out = []
    for comp in comps:
        names = sorted(NA.get_base_atom_names(comp.sequence))
        centers = {n: [float(v) for v in comp.centers[n]] for n in names if len(comp.centers[n]) == 3}
        points = [[a.name, float(a.x), float(a.y), float(a.z)] for a in comp.atoms() if a.name in set(names)]
        out.append({"unit_id": comp.unit_id(), "sequence": comp.sequence,
                    "center": [float(v) for v in comp.centers["base"]],
                    "rotation": np.asarray(comp.rotation_matrix).tolist(),
                    "centers": centers, "base_points": points})
    return out


def frames_golden(spec):
    comps = residues(rf.ref_structure_from_spec(spec))
    observed = [set(a[0] for a in nt["atoms"]) for nt in spec["nts"]]
    out = []
    for comp, obs in zip(comps, observed):
        hyd = {a.name: [float(a.x), float(a.y), float(a.z)] for a in comp.atoms() if a.name not in obs}
        out.append({"unit_id": comp.unit_id(),
                    "center": [float(v) for v in comp.centers["base"]],
                    "rotation": np.asarray(comp.rotation_matrix).tolist(),
                    "hydrogens": hyd})
    return out


def annotate_with_traces(spec, categories):
    """Run annotate_nt_nt_in_structure with every rule function wrapped so each
    call's (nt1, nt2, result) is recorded in call order."""
    from fr3d.classifiers import NA_pairwise_interactions as NA
    structure = rf.ref_structure_from_spec(spec)
    comps = residues(structure)
    idx = {c.unit_id(): k for k, c in enumerate(comps)}
    tr = {"candidates": [], "sO": [], "stacking": [], "sugar_ribose": [], "backbone": [],
          "coplanar": [], "basepair": []}
    orig = {}

    def wrap(name, fn):
        orig[name] = getattr(NA, name)
        setattr(NA, name, fn)

    def so(nt1, nt2, parent1, dp):
        r = orig["check_base_oxygen_stack_rings"](nt1, nt2, parent1, dp)
        tr["sO"].append([idx[nt1.unit_id()], idx[nt2.unit_id()], r[0]])
        return r

    def stack(nt1, nt2, p1, p2, dp):
        r = orig["check_base_base_stacking"](nt1, nt2, p1, p2, dp)
        tr["stacking"].append([idx[nt1.unit_id()], idx[nt2.unit_id()], r[0]])
        return r

    def sr(nt1, nt2, parent1, dp):
        r = orig["check_sugar_ribose"](nt1, nt2, parent1, dp)
        tr["sugar_ribose"].append([idx[nt1.unit_id()], idx[nt2.unit_id()], r[0]])
        return r

    def bb(nt1, nt2, prevO3, p1, p2, dp):
        r = orig["check_base_backbone_interactions"](nt1, nt2, prevO3, p1, p2, dp)
        tr["backbone"].append([idx[nt1.unit_id()], idx[nt2.unit_id()], r[0], r[1]])
        return r

    def cp(nt1, nt2, pair_data, dp):
        r = orig["check_coplanar"](nt1, nt2, pair_data, dp)
        pd = r[0]
        tr["coplanar"].append([idx[nt1.unit_id()], idx[nt2.unit_id()], bool(pd["coplanar"]),
                               float(pd["gap12"]),
                               float(pd["min_distance"]) if "min_distance" in pd else None,
                               float(pd["coplanar_value"]) if pd["coplanar_value"] is not None else None])
        return r

    def bp(nt1, nt2, pair_data, cutoffs, hb, dp):
        r = orig["check_basepair_cutoffs"](nt1, nt2, pair_data, cutoffs, hb, dp)
        q = r[2]
        tr["basepair"].append([idx[nt1.unit_id()], idx[nt2.unit_id()], r[0],
                               r[1] if r[0] else None,
                               float(q["cutoff_distance"]) if q else None,
                               float(q["max_gap"]) if q else None])
        return r

    # candidate order: the screens end right before parent2 = get_parent(nt2.sequence)
    # (NA_pairwise_interactions.py:726-729); unit_id_pair is built first.  Wrap the
    # first per-candidate call instead: with 'sO' on it is check_base_oxygen_stack_rings(nt1,nt2).
    wrap("check_base_oxygen_stack_rings", so)
    wrap("check_base_base_stacking", stack)
    wrap("check_sugar_ribose", sr)
    wrap("check_base_backbone_interactions", bb)
    wrap("check_coplanar", cp)
    wrap("check_basepair_cutoffs", bp)
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            result = NA.annotate_nt_nt_in_structure(structure, categories)
    finally:
        for name, fn in orig.items():
            setattr(NA, name, fn)
    if "stacking" in categories:
        tr["candidates"] = [[a, b] for a, b, _ in tr["stacking"]]
    i2t = {label: [[idx[a], idx[b], c] for a, b, c in lst] for label, lst in result[0].items()}
    c2i = {cat: sorted(v) for cat, v in result[1].items()}
    found = [l for l in buf.getvalue().splitlines() if "Found" in l]
    return {"categories": categories, "unit_ids": [c.unit_id() for c in comps],
            "traces": tr, "interaction_to_list_of_tuples": i2t,
            "category_to_interactions": c2i, "found_line": found[0] if found else ""}


def geometry_golden(strand_spec, sarcin_spec):
    from fr3d.geometry.superpositions import besttransformation, besttransformation_weighted
    from fr3d.geometry.discrepancy import matrix_discrepancy, matrix_discrepancy_cutoff
    from fr3d.search import discrepancy as sdisc
    from fr3d.geometry.angleofrotation import angle_of_rotation
    out = {}

    # --- Kabsch KATs: inputs of tests/superpositions_test.py:11-119, outputs from the reference
    a = np.array([[1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, -1, 0], [0, 0, 1.0], [0, 0, -1.0]])
    b = np.array([[0, 1, 0], [-1, 0, 0], [0, -1, 0], [1, 0, 0], [0, 0, 1.0], [0, 0, -1.0]])
    kats = [("octahedron_90z", a, b)]
    p4 = np.array([[2.0, 0, 0], [0, 3.0, 0], [0, 0, 4.0], [1.0, 1.0, 1.0]])

    def rot(axis, ang):
        c, s = np.cos(ang), np.sin(ang)
        if axis == 'x':
            return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
        if axis == 'y':
            return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])

    for ax in "xyz":
        kats.append(("p4_45" + ax, p4, p4 @ rot(ax, np.pi / 4).T))
    kats.append(("p4_xyz", p4, p4 @ (rot('x', 0.3) @ rot('y', -0.7) @ rot('z', 1.1)).T))
    # reflection case: best fit needs the det fix
    refl = p4 * np.array([1, 1, -1.0])
    kats.append(("p4_reflected", p4, refl))
    rng = np.random.Generator(np.random.PCG64(77))
    for k in range(40):
        n = int(rng.integers(3, 12))
        s1 = rng.normal(size=(n, 3)) * 5
        R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        s2 = s1 @ R + rng.normal(size=(n, 3)) * (0.0 if k % 4 == 0 else 0.3) + rng.normal(size=3) * 10
        kats.append(("rand%02d" % k, s1, s2))
    kab = []
    for name, s1, s2 in kats:
        U, new1, mean1, rmsd, sse, mean2 = besttransformation(s1, s2)
        w = (1.0 + np.arange(len(s1)) % 3).tolist()
        Uw, new1w, mean1w, rmsdw, ssew = besttransformation_weighted(s1, s2, w)
        kab.append({"name": name, "set1": s1.tolist(), "set2": s2.tolist(),
                    "U": np.asarray(U).tolist(), "new1": np.asarray(new1).tolist(),
                    "mean1": np.asarray(mean1).tolist(), "mean2": np.asarray(mean2).tolist(),
                    "rmsd": float(rmsd), "sse": float(sse),
                    "weights": w, "Uw": np.asarray(Uw).tolist(), "ssew": float(ssew), "rmsdw": float(rmsdw),
                    "angle": float(angle_of_rotation(np.asarray(U)))})
    out["kabsch"] = kab

    # --- discrepancy KATs (tests/discrepancy_tests.py:301-335)
    st = residues(rf.ref_structure_from_spec(strand_spec))
    by = {int(c.number): c for c in st}

    def info(nums, source):
        return ([np.asarray(source[k].centers["base"]) for k in nums],
                [np.asarray(source[k].rotation_matrix) for k in nums])

    dk = []
    for name, n1, n2 in [("simple", [10, 11, 12, 13, 14], [15, 16, 17, 18, 19]),
                         ("another", [10, 12, 14, 16, 18], [11, 13, 15, 17, 19]),
                         ("pair2", [10, 12], [11, 13])]:
        c1, r1 = info(n1, by)
        c2, r2 = info(n2, by)
        d = float(matrix_discrepancy(c1, r1, c2, r2))
        ds = float(sdisc.matrix_discrepancy(c1, r1, c2, r2))
        dc_hi = matrix_discrepancy_cutoff(c1, r1, c2, r2, 2.5)
        dc_lo = matrix_discrepancy_cutoff(c1, r1, c2, r2, 0.5)
        dk.append({"name": name, "nts1": n1, "nts2": n2, "d": d, "d_search_copy": ds,
                   "cutoff_2.5": None if dc_hi is None else float(dc_hi),
                   "cutoff_0.5": None if dc_lo is None else float(dc_lo),
                   "centers1": np.array(c1).tolist(), "rotations1": np.array(r1).tolist(),
                   "centers2": np.array(c2).tolist(), "rotations2": np.array(r2).tolist()})
    sa = residues(rf.ref_structure_from_spec(sarcin_spec))
    c1 = [np.asarray(c.centers["base"]) for c in sa[:7]]
    r1 = [np.asarray(c.rotation_matrix) for c in sa[:7]]
    c2 = [np.asarray(c.centers["base"]) for c in sa[7:]]
    r2 = [np.asarray(c.rotation_matrix) for c in sa[7:]]
    dk.append({"name": "sarcin7", "d": float(matrix_discrepancy(c1, r1, c2, r2)),
               "centers1": np.array(c1).tolist(), "rotations1": np.array(r1).tolist(),
               "centers2": np.array(c2).tolist(), "rotations2": np.array(r2).tolist()})
    sel = [0, 1, 2, 4, 5]
    dk.append({"name": "sarcin5", "d": float(matrix_discrepancy([c1[k] for k in sel], [r1[k] for k in sel],
                                                                 [c2[k] for k in sel], [r2[k] for k in sel])),
               "centers1": np.array([c1[k] for k in sel]).tolist(), "rotations1": np.array([r1[k] for k in sel]).tolist(),
               "centers2": np.array([c2[k] for k in sel]).tolist(), "rotations2": np.array([r2[k] for k in sel]).tolist()})
    out["discrepancy_kats"] = dk

    # --- a seeded all-vs-all block: instance set made like SURVEY §8(d) C5 (seed 5000), N=48, n=7
    inst_c, inst_r = make_instances(np.array(c1), np.array(r1), 48, 5000)
    N = len(inst_c)
    D = np.zeros((N, N))
    for i in range(N):
        for j in range(i + 1, N):
            D[i, j] = matrix_discrepancy(list(inst_c[i]), list(inst_r[i]), list(inst_c[j]), list(inst_r[j]))
            D[j, i] = D[i, j]
    # position-0 rotation-less variant (NA_protein_annotation.py:1705-1714)
    Dn = np.zeros((N, N))
    empty = np.empty((0, 0))
    for i in range(N):
        for j in range(i + 1, N):
            ri = [empty] + list(inst_r[i][1:])
            rj = [empty] + list(inst_r[j][1:])
            Dn[i, j] = matrix_discrepancy(list(inst_c[i]), ri, list(inst_c[j]), rj)
            Dn[j, i] = Dn[i, j]
    cut = 0.6
    row0 = [matrix_discrepancy_cutoff(list(inst_c[0]), list(inst_r[0]), list(inst_c[j]), list(inst_r[j]), cut)
            for j in range(N)]
    out["all_vs_all"] = {"seed": 5000, "N": N, "n": 7, "base_centers": np.array(c1).tolist(),
                         "base_rotations": np.array(r1).tolist(),
                         "centers": inst_c.tolist(), "rotations": inst_r.tolist(),
                         "D": D.tolist(), "D_norot0": Dn.tolist(), "cutoff": cut,
                         "row0_cutoff": [None if v is None else float(v) for v in row0]}
    return out


def make_instances(base_c, base_r, N, seed):
    """SURVEY §8(d) C5 recipe: rigid motion + centre noise N(0,0.5) + per-rotation
    perturbation of angle ~N(0,0.15 rad).  MUST stay in sync with
    fr3d_python_b200.synthetic.make_instances (tests compare the arrays)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = base_c.shape[0]
    C = np.zeros((N, n, 3))
    R = np.zeros((N, n, 3, 3))
    for k in range(N):
        Q = _random_rotation(rng)
        t = rng.normal(size=3) * 20.0
        C[k] = base_c @ Q.T + t + rng.normal(size=(n, 3)) * 0.5
        for p in range(n):
            axis = rng.normal(size=3)
            axis /= np.linalg.norm(axis)
            ang = rng.normal() * 0.15
            R[k, p] = Q @ base_r[p] @ _axis_angle(axis, ang)
    return C, R


def _random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _axis_angle(axis, ang):
    x, y, z = axis
    K = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)


ALL9 = {'basepair': LW18 + ['cWB', 'cBW'], 'basepair_detail': [], 'stacking': [], 'sO': [], 'backbone': [],
        'coplanar': [], 'sugar_ribose': [], 'covalent': [], 'near': []}


def main():
    os.makedirs(GOLD, exist_ok=True)
    base = rf.spec_1gid_base()
    strand = rf.spec_1s72_strand()
    sarcin = rf.spec_1s72_sarcin()
    full = make_spec_1gid_full(base, strand)
    dump("spec_1gid_base.json.gz", base)
    dump("spec_1gid_full.json.gz", full)
    dump("spec_1s72_strand.json.gz", strand)
    dump("spec_1s72_sarcin.json.gz", sarcin)
    dump("matlab_1gid_frames.json.gz", rf.parse_matlab_rotations())
    for name, spec in [("1gid_full", full), ("1s72_strand", strand), ("1s72_sarcin", sarcin)]:
        dump("frames_%s.json.gz" % name, frames_golden(spec))
    dump("annot_1gid_full_all9.json.gz", annotate_with_traces(full, ALL9))
    dump("annot_1s72_strand_all9.json.gz", annotate_with_traces(strand, ALL9))
    dump("annot_1gid_base_bp_stack_cp.json.gz",
         annotate_with_traces(base, {'basepair': LW18, 'stacking': [], 'coplanar': []}))
    dump("annot_1gid_base_bp.json.gz", annotate_with_traces(base, {'basepair': LW18}))
    dump("geometry.json.gz", geometry_golden(strand, sarcin))
    modified = make_spec_modified(full)
    dump("spec_1gid_modified.json.gz", modified)
    dump("frames_1gid_modified.json.gz", frames_golden_modified(modified))
    dump("annot_1gid_modified_all9.json.gz", annotate_with_traces(modified, ALL9))


if __name__ == "__main__" and len(sys.argv) == 1:
    main()


def make_spec_variants(full_spec):
    """1GID-full with the identity / completeness corner cases the reference handles on its way to the rules:
    DNA residues (DA DC DG DT, no O2'), missing backbone / glycosidic / base atoms, alt-id twins (NA:706-713),
    an insertion-code pair, a second model (separate cubes, NA:424-436) and a symmetry copy (same cubes)."""
    import copy
    spec = copy.deepcopy(full_spec)
    comps = residues(rf.ref_structure_from_spec(full_spec))
    from fr3d.definitions import NAbasecoordinates
    nts = spec["nts"]
    for k, nt in enumerate(nts[:120]):
        if k % 3 == 0 and nt["sequence"] in ("A", "C", "G"):
            nt["sequence"] = "D" + nt["sequence"]
            nt["atoms"] = [a for a in nt["atoms"] if a[0] != "O2'"]
        elif k % 3 == 1 and nt["sequence"] == "U":
            nt["sequence"] = "DT"
            U = np.asarray(comps[k].rotation_matrix)
            c7 = np.asarray(comps[k].centers["base"]) + U @ np.array(NAbasecoordinates["DT"]["C7"])
            nt["atoms"] = [a for a in nt["atoms"] if a[0] != "O2'"] + [["C7", float(c7[0]), float(c7[1]), float(c7[2])]]
        elif k % 7 == 2:
            nt["atoms"] = [a for a in nt["atoms"] if a[0] not in ("P", "OP1", "O4'", "C2'")]
        elif k % 11 == 5:
            nt["atoms"] = [a for a in nt["atoms"] if a[0] not in ("N9", "N1")]
        elif k % 13 == 6:
            nt["atoms"] = [a for a in nt["atoms"] if a[0] not in ("N3", "O2", "C5")]
    out = []
    for k, nt in enumerate(nts):
        if k in (130, 131, 200):                       # alt-id twins: same residue, conformers A and B
            a = copy.deepcopy(nt); a["alt_id"] = "A"
            b = copy.deepcopy(nt); b["alt_id"] = "B"
            b["atoms"] = [[n, x + 0.25, y - 0.15, z + 0.2] for n, x, y, z in b["atoms"]]
            out += [a, b]
        elif k == 150:                                 # insertion codes: 150A follows 150
            a = copy.deepcopy(nt)
            out.append(a)
        elif k == 151:
            a = copy.deepcopy(nt)
            a["number"] = nts[150]["number"]
            a["insertion_code"] = "A"
            out.append(a)
        else:
            out.append(nt)
    first_chain = nts[0]["chain"]
    model2 = []
    for nt in nts[:40]:                                # second model: same place, must not pair with model 1
        a = copy.deepcopy(nt)
        a["model"] = "2"
        a["atoms"] = [[n, x + 0.4, y, z] for n, x, y, z in a["atoms"]]
        model2.append(a)
    sym = []
    for nt in nts[60:90]:                              # symmetry copy next to the original: pairs across copies
        a = copy.deepcopy(nt)
        a["symmetry"] = "2_655"
        a["atoms"] = [[n, x + 9.0, y + 4.0, z - 6.0] for n, x, y, z in a["atoms"]]
        sym.append(a)
    assert first_chain
    return {"pdb": "1GID", "nts": out + model2 + sym}


def main_variants():
    with gzip.open(os.path.join(GOLD, "spec_1gid_full.json.gz"), "rt") as f:
        full = json.load(f)
    variants = make_spec_variants(full)
    dump("spec_1gid_variants.json.gz", variants)
    dump("frames_1gid_variants.json.gz", frames_golden(variants))
    dump("annot_1gid_variants_all9.json.gz", annotate_with_traces(variants, ALL9))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "variants":
    main_variants()


def main_category_subsets():
    """Category subsets ('basepair' must always be a key: the reference indexes categories['basepair'], NA:1216) and
    focused interaction lists (focus_basepair_cutoffs keeps only the listed families, NA:88-136)."""
    with gzip.open(os.path.join(GOLD, "spec_1gid_full.json.gz"), "rt") as f:
        full = json.load(f)
    out = {}
    for name, cats in [("detail", {'basepair': [], 'sO': [], 'backbone': [], 'sugar_ribose': []}),
                       ("focus_cww_ths_stacking", {'basepair': ['cWW', 'tHS'], 'stacking': []}),
                       ("focus_cww_coplanar_covalent", {'basepair': ['cWW'], 'coplanar': [], 'covalent': []})]:
        g = annotate_with_traces(full, cats)
        out[name] = {"categories": cats, "interaction_to_list_of_tuples": g["interaction_to_list_of_tuples"],
                     "category_to_interactions": g["category_to_interactions"], "found_line": g["found_line"]}
    out["unit_ids"] = g["unit_ids"]
    dump("annot_1gid_full_category_subsets.json.gz", out)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "subsets":
    main_category_subsets()


def inflated_cutoffs(NA, only_largest):
    """The table of tests/test_gpu_parity.py::_inflated_cutoffs, built from the reference's own tables."""
    import copy
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


def main_custom_cutoffs():
    """Caller-supplied focused_basepair_cutoffs (third argument of annotate_nt_nt_in_structure)."""
    from fr3d.classifiers import NA_pairwise_interactions as NA
    with gzip.open(os.path.join(GOLD, "spec_1gid_full.json.gz"), "rt") as f:
        full = json.load(f)
    out = {}
    for only_largest in (True, False):
        structure = rf.ref_structure_from_spec(full)
        comps = residues(structure)
        idx = {c.unit_id(): k for k, c in enumerate(comps)}
        focused = inflated_cutoffs(NA, only_largest)
        with contextlib.redirect_stdout(io.StringIO()):
            result = NA.annotate_nt_nt_in_structure(structure, ALL9, focused)
        out["only_largest" if only_largest else "all"] = {
            "interaction_to_list_of_tuples": {label: [[idx[a], idx[b], c] for a, b, c in lst]
                                              for label, lst in result[0].items()},
            "category_to_interactions": {cat: sorted(v) for cat, v in result[1].items()}}
        out["unit_ids"] = [c.unit_id() for c in comps]
    dump("annot_1gid_full_custom_cutoffs.json.gz", out)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "cutoffs":
    main_custom_cutoffs()


def reference_bond_orientation():
    """annotate_bond_orientation of fr3d/classifiers/NA_unit_annotation.py, UNMODIFIED.  The module cannot be imported
    (its top imports fr3d.cif.reader -> pdbx, and bare `from NA_pairwise_interactions import ...`), so the function's
    own source lines are compiled from the reference file with the four names it reads bound to the reference's
    objects."""
    import ast
    import math
    from fr3d.classifiers import NA_pairwise_interactions as NA
    from fr3d.data.mapping import parent_atom_to_modified
    path = os.path.join(rf.REF, "fr3d", "classifiers", "NA_unit_annotation.py")
    with open(path) as f:
        tree = ast.parse(f.read(), path)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "annotate_bond_orientation"]
    mod = ast.Module(body=fn, type_ignores=[])
    env = {"np": np, "math": math, "get_parent": NA.get_parent, "parent_atom_to_modified": parent_atom_to_modified}
    exec(compile(mod, path, "exec"), env)
    return env["annotate_bond_orientation"]


def make_spec_chi_sweep(strand_spec):
    """Four nts of the 1S72 strand (A C G U) with the base turned about the glycosidic bond in 15-degree steps, so
    that every orientation class (anti, syn, int_syn) and the class boundaries' neighbourhoods occur."""
    nts = []
    picked = {}
    for nt in strand_spec["nts"]:
        picked.setdefault(nt["sequence"], nt)
    k = 0
    for seq in ("A", "C", "G", "U"):
        nt = picked[seq]
        xyz = {a[0]: np.array(a[1:], dtype=float) for a in nt["atoms"]}
        n_name = "N9" if seq in ("A", "G") else "N1"
        axis = xyz[n_name] - xyz["C1'"]
        axis /= np.linalg.norm(axis)
        for step in range(24):
            R = _axis_angle(axis, np.deg2rad(15.0 * step))
            atoms = []
            for name, x, y, z in nt["atoms"]:
                p = np.array([x, y, z], dtype=float)
                if name not in BACKBONE:
                    p = xyz[n_name] + R.dot(p - xyz[n_name])
                atoms.append([name] + [float(v) for v in p + np.array([40.0 * k, 0.0, 0.0])])
            k += 1
            nts.append({"chain": "S", "sequence": seq, "number": str(k), "index": k, "model": nt["model"],
                        "symmetry": nt["symmetry"], "alt_id": None, "insertion_code": None, "atoms": atoms})
    return {"pdb": "CHI", "nts": nts}


def main_glycosidic():
    """chi / syn-anti annotation: the reference's rows on the full, modified and variant structures
    and on a sweep of the glycosidic torsion."""
    fn = reference_bond_orientation()
    out = {}
    with gzip.open(os.path.join(GOLD, "spec_1s72_strand.json.gz"), "rt") as f:
        dump("spec_chi_sweep.json.gz", make_spec_chi_sweep(json.load(f)))
    for name in ("1gid_full", "1gid_modified", "1gid_variants", "1s72_strand", "chi_sweep"):
        with gzip.open(os.path.join(GOLD, "spec_%s.json.gz" % name), "rt") as f:
            spec = json.load(f)
        structure = rf.ref_structure_from_spec(spec)
        with contextlib.redirect_stdout(io.StringIO()) as printed:
            rows, errors = fn(structure, pipeline=True)
        out[name] = {"rows": rows, "errors": errors, "printed": printed.getvalue()}
    dump("glycosidic.json.gz", out)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "glycosidic":
    main_glycosidic()


def main_search_screen():
    """FR3D search screening: fixed_radius_search and get_pairlist of
    fr3d/search/pair_processing.py (imported unmodified: it only needs numpy) on the base centres of the full and the
    variants structures (two models, some points without a centre)."""
    sys.path.insert(0, os.path.join(rf.REF, "fr3d", "search"))
    import pair_processing as pp
    out = {}
    for name in ("1gid_full", "1gid_variants"):
        with gzip.open(os.path.join(GOLD, "frames_%s.json.gz" % name), "rt") as f:
            frames = json.load(f)
        with gzip.open(os.path.join(GOLD, "spec_%s.json.gz" % name), "rt") as f:
            spec = json.load(f)
        centers = np.array([fr["center"] for fr in frames], dtype=float)
        models = [str(nt["model"]) for nt in spec["nts"]]
        if name == "1gid_variants":
            centers[5::41] = np.nan                      # units without a centre (pair_processing.py:137)
            centers -= centers[np.isfinite(centers[:, 0])].mean(axis=0) * 1.7      # negative coordinates too
        case = {"centers": [[None if np.isnan(v) else float(v) for v in c] for c in centers], "models": models,
                "searches": {}}
        mn, mx = np.nanmin(centers), np.nanmax(centers)
        square_size = float(np.power(10, np.floor(np.log10(mx - mn))) * np.ceil((mx - mn) / np.power(10, np.floor(np.log10(mx - mn)))))
        for cutoff in ((6.0, 12.0, 21.5) if name == "1gid_full" else (6.0, 12.0)):
            pairs = pp.fixed_radius_search(len(centers), models, centers, cutoff, square_size)
            case["searches"][str(cutoff)] = [[int(a), int(b), float(c)] for a, b, c in pairs]
        Q = {"type": "geometric", "numpositions": 3, "largestMaxRange": 14.0,
             "requireddistanceminimum": [[0, 0.0, 4.0], [0.0, 0, 7.5], [4.0, 7.5, 0]],
             "requireddistancemaximum": [[0, 9.0, 14.0], [9.0, 0, 12.25], [14.0, 12.25, 0]]}
        pl = pp.get_pairlist(Q, models, centers)
        case["Q"] = Q
        case["pairlist"] = {"%d,%d" % (i, j): [[int(a), int(b)] for a, b in v] for i, d in pl.items() for j, v in d.items()}
        out[name] = case
    dump("search_screen.json.gz", out)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "search":
    main_search_screen()


# ---------------------------------------------------------------------------------------------------------------
# mmCIF ingest: what the UNMODIFIED fr3d.cif.reader makes of a CIF text.  The reference needs the
# `pdbx` package, which is not installed: tools/pdbx_shim provides the few classes it uses (dev container only).
def make_cif_scenario():
    """A small mmCIF file with the cases the reader's residue delimitation has to get right (reader.py:440-629):
    DNA residues, missing atoms, modified residues, alternate conformers with common atoms, an insertion code, a
    second model, two extra symmetry operators on one chain, hetero / water / peptide residues to be filtered out."""
    import copy
    with gzip.open(os.path.join(GOLD, "spec_1gid_variants.json.gz"), "rt") as f:
        variants = json.load(f)
    with gzip.open(os.path.join(GOLD, "spec_1gid_modified.json.gz"), "rt") as f:
        modified = json.load(f)
    pool = [nt for nt in variants["nts"] if nt["symmetry"] == "1_555"]
    m1 = [nt for nt in pool if nt["model"] == "1"]
    keep = m1[:40] + [nt for nt in m1 if nt["chain"] == "A" and 226 <= int(nt["number"]) <= 256]
    keep += [nt for nt in m1 if nt["chain"] == "B"][:14]
    keep += [nt for nt in pool if nt["model"] == "2"][:10]
    nts = copy.deepcopy(keep)
    mod = {(nt["chain"], nt["number"]): nt for nt in modified["nts"] if nt["sequence"] not in "ACGU"}
    swapped = 0
    for k, nt in enumerate(nts):                   # a few modified residues (PSU, 5MC, OMG, ...)
        key = (nt["chain"], nt["number"])
        if key in mod and nt["model"] == "1" and not nt.get("alt_id") and swapped < 4 and nt["sequence"] in "ACGU":
            nts[k] = copy.deepcopy(mod[key])
            swapped += 1
    # alternate conformers: the phosphate of conformer A becomes common to both (alt '.'), conformer B loses it
    for nt in nts:
        if nt.get("alt_id") == "A":
            nt["atom_alt"] = [None if a[0] in ("P", "OP1", "OP2") else "A" for a in nt["atoms"]]
        elif nt.get("alt_id") == "B":
            nt["atoms"] = [a for a in nt["atoms"] if a[0] not in ("P", "OP1", "OP2")]
    for nt in nts:
        nt["atoms"] = [[a[0], round(a[1], 3), round(a[2], 3), round(a[3], 3)] for a in nt["atoms"]]
    hetero = [
        {"chain": "A", "sequence": "MG", "number": "901", "index": None, "model": "1", "group": "HETATM", "entity": 2,
         "atoms": [["MG", 52.104, 48.331, 87.25]]},
        {"chain": "A", "sequence": "HOH", "number": "902", "index": None, "model": "1", "group": "HETATM", "entity": 3,
         "atoms": [["O", 50.0, 47.5, 86.125]]},
        {"chain": "C", "sequence": "ALA", "number": "5", "index": 5, "model": "1", "group": "ATOM", "entity": 4,
         "atoms": [["N", 40.0, 40.5, 80.25], ["CA", 41.2, 40.9, 80.75], ["C", 42.0, 39.8, 81.5], ["O", 41.7, 38.6, 81.375]]},
    ]
    c, s = np.cos(np.pi / 6), np.sin(np.pi / 6)
    operators = [
        {"id": "1", "name": "1_555", "matrix": [[1, 0, 0], [0, 1, 0], [0, 0, 1]], "vector": [0, 0, 0]},
        {"id": "2", "name": "2_655", "matrix": [[-1, 0, 0], [0, -1, 0], [0, 0, 1]], "vector": [104.5, 110.25, 3.5]},
        {"id": "3", "name": "?", "matrix": [[c, -s, 0], [s, c, 0], [0, 0, 1]], "vector": [12.125, -7.5, 20.0]},
    ]
    assembly = [("1,2,3", "A"), ("1", "B,C")]
    return {"pdb": "1GID", "nts": nts}, hetero, operators, assembly


def main_cif():
    sys.path.insert(0, os.path.join(HERE, "pdbx_shim"))
    sys.path.insert(0, os.path.dirname(HERE))
    import contextlib
    import io
    from fr3d.cif.reader import Cif
    from fr3d_python_b200.cif import writer
    spec, hetero, operators, assembly = make_cif_scenario()
    text = writer.spec_to_cif(spec, operators=operators, assembly=assembly, hetero=hetero,
                              extra_chem={"MG": "non-polymer", "HOH": "non-polymer", "ALA": "L-peptide linking"})
    with contextlib.redirect_stdout(io.StringIO()):
        st = Cif(io.StringIO(text)).structure()
    out = []
    for r in residues(st):
        # the atoms the reader read: Component.__init__ appends inferred hydrogens (and, for modified residues without
        # hydrogens, ideal-position copies of every mapped base atom) AFTER them - keep the first atom of each name
        seen, atoms = set(), []
        for a in r.atoms():
            if a.name.startswith("H") or a.name in seen:
                continue
            seen.add(a.name)
            atoms.append([a.name, float(a.x), float(a.y), float(a.z)])
        out.append({"unit_id": r.unit_id(), "chain": r.chain, "sequence": r.sequence, "number": r.number, "index": r.index,
                    "model": r.model, "symmetry": r.symmetry, "alt_id": r.alt_id, "insertion_code": r.insertion_code,
                    "atoms": atoms})
    with gzip.GzipFile(os.path.join(GOLD, "cif_scenario.cif.gz"), "wb", mtime=0) as f:
        f.write(text.encode())
    print("wrote cif_scenario.cif.gz (%d bytes of text)" % len(text))
    dump("cif_scenario_reference_reader.json.gz", {"pdb": st.pdb, "nts": out})
    print("residues: %d nucleotides; symmetries %s; models %s" % (len(out), sorted(set(o["symmetry"] for o in out)),
                                                                  sorted(set(o["model"] for o in out))))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "cif":
    main_cif()


# ---------------------------------------------------------------------------------------------------------------
# Ordering by similarity (last part): the UNMODIFIED fr3d/ordering/orderBySimilarity.py and
# fr3d/search/orderBySimilarity.py (the copy search/FR3D.py:463 calls) on seeded matrices.
def ordering_matrices():
    rng = np.random.Generator(np.random.PCG64(8165))
    mats = {}
    for n in (3, 4, 5, 6, 7, 17, 60, 150):
        pts = rng.uniform(0, 1, (n, 3))
        if n >= 60:                                   # clustered points: the case the penalty is made for
            pts = rng.uniform(0, 1, (6, 3))[rng.integers(0, 6, n)] + rng.normal(0, 0.05, (n, 3))
        mats["points_%d" % n] = np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1))
    for n in (12, 40):                                # tie-laden: L1 distances of small-integer grid points
        pts = rng.integers(0, 4, (n, 2)).astype(float)
        mats["ties_%d" % n] = np.abs(pts[:, None, :] - pts[None, :, :]).sum(-1)
    mats["zeros_5"] = np.zeros((5, 5))
    # discrepancies of perturbed sarcin-ricin instances (what FR3D.py:433-442 feeds the ordering)
    from fr3d.geometry.discrepancy import matrix_discrepancy
    with gzip.open(os.path.join(GOLD, "frames_1s72_sarcin.json.gz"), "rt") as f:
        frames = json.load(f)
    base_c = np.array([fr["center"] for fr in frames][:7])
    base_r = np.array([fr["rotation"] for fr in frames][:7])
    cen, rot = make_instances(base_c, base_r, 40, 5001)
    with contextlib.redirect_stdout(io.StringIO()):
        d = np.zeros((40, 40))
        for i in range(40):
            for j in range(i + 1, 40):
                d[i][j] = d[j][i] = matrix_discrepancy(list(cen[i]), list(rot[i]), list(cen[j]), list(rot[j]))
    mats["sarcin_40"] = d
    for k in mats:
        np.fill_diagonal(mats[k], 0.0)
        mats[k] = (mats[k] + mats[k].T) / 2
    return mats


def main_ordering():
    import importlib.util
    import random
    from fr3d.ordering import orderBySimilarity as obs
    sys.path.insert(0, os.path.join(rf.REF, "fr3d", "search"))
    spec = importlib.util.spec_from_file_location("search_orderBySimilarity",
                                                  os.path.join(rf.REF, "fr3d", "search", "orderBySimilarity.py"))
    sobs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sobs)
    out = {}
    for name, d in ordering_matrices().items():
        n = d.shape[0]
        case = {"distance": d.tolist()}
        pen = obs.treePenalty(d)
        assert np.array_equal(pen, sobs.treePenalty(d))
        case["penalty"] = pen.tolist()
        case["normalized"] = np.asarray(obs.normalizePenaltyMatrix(d, pen)).tolist()
        random.seed(n + 100)
        start = list(range(n))
        random.shuffle(start)
        path, score = obs.greedyInsertionPathLength(d, list(start))
        swapped, reduction = obs.twoOptSwap(d, list(path))
        case["start"] = start
        case["greedy"] = [[int(v) for v in path], float(score)]
        case["twoopt"] = [[int(v) for v in swapped], float(reduction)]
        swapped2, reduction2 = obs.twoOptSwap(d, list(start))          # 2-opt from a random order: many swaps
        case["twoopt_from_start"] = [[int(v) for v in swapped2], float(reduction2)]
        case["tppl"] = {}
        for reps, seed, strength in ((10, 1, 0.5), (5, 7, 1.0), (3, 2, 0.0)):
            order = obs.treePenalizedPathLength(d, reps, seed, strength)
            case["tppl"]["%d,%d,%g" % (reps, seed, strength)] = {
                "order": [int(v) for v in order],
                "standard": [int(v) for v in obs.standardOrder(d, list(order))]}
        case["tppl_search"] = {}
        for reps, seed in ((100, 59), (7, 3)):
            case["tppl_search"]["%d,%d" % (reps, seed)] = [int(v) for v in sobs.treePenalizedPathLength(d, reps, seed)]
        case["reordered"] = np.asarray(obs.reorderSymmetricMatrix(d, case["tppl"]["10,1,0.5"]["order"])).tolist()
        out[name] = case
    dump("ordering.json.gz", out)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "ordering":
    main_ordering()


# ---------------------------------------------------------------------------------------------------------------
# FR3D search, constraint intersection (search.py:263-398).  fr3d/search/search.py is imported UNMODIFIED; its
# `from query_processing import synonym` cannot work (query_processing.py has an IndentationError under Python 3), so that
# one module is stubbed - none of the functions used here touches it.
def import_reference_search():
    import types
    sys.path.insert(0, os.path.join(rf.REF, "fr3d", "search"))
    if "query_processing" not in sys.modules:
        stub = types.ModuleType("query_processing")
        stub.synonym = {}
        sys.modules["query_processing"] = stub
    import search as ref_search
    return ref_search


def possibility_scenarios():
    """Inputs of the intersection: (Q, listOfPairs, universe, centers) per scenario, before reorderPositions."""
    import pair_processing as pp
    with gzip.open(os.path.join(GOLD, "frames_1gid_full.json.gz"), "rt") as f:
        frames = json.load(f)
    centers = np.array([fr["center"] for fr in frames], dtype=float)
    models = ["1"] * len(centers)
    n_units = len(centers)
    out = {}
    # geometric queries: a unit and its nearest neighbours as the query motif (query_processing.py:886-914, :1113-1121)
    for name, seed_unit, k, disc in (("geometric_4", 20, 4, 0.6), ("geometric_5", 150, 5, 0.5), ("geometric_3", 77, 3, 0.8),
                                     ("geometric_2", 5, 2, 0.5)):
        d = np.linalg.norm(centers - centers[seed_unit], axis=1)
        motif = [int(v) for v in np.argsort(d)[:k]]
        Q = {"type": "geometric", "numpositions": k, "discrepancy": disc, "locationWeight": [1.0] * k,
             "centers": [centers[u] for u in motif]}
        Q["requireddistanceminimum"] = {i: {} for i in range(k)}
        Q["requireddistancemaximum"] = {i: {} for i in range(k)}
        Q["SSCutoff"] = [None] * k
        Q["largestMaxRange"] = 0
        Q["distance"] = np.zeros((k, k))
        s = 0
        for i in range(k):
            s = s + Q["locationWeight"][i]
            Q["SSCutoff"][i] = (k ** 2) * (disc ** 2) * s
            for j in range(i + 1, k):
                Q["distance"][i][j] = Q["distance"][j][i] = np.linalg.norm(Q["centers"][i] - Q["centers"][j])
                delta = (np.sqrt(2.0) if k > 2 else 1.0) * k * disc
                for a, b in ((i, j), (j, i)):
                    Q["requireddistanceminimum"][a][b] = Q["distance"][i][j] - delta
                    Q["requireddistancemaximum"][a][b] = Q["distance"][i][j] + delta
                Q["largestMaxRange"] = max(Q["largestMaxRange"], Q["distance"][i][j] + delta)
        Q["cutoff"] = [float(v) for v in Q["SSCutoff"]]
        pl = pp.get_pairlist(Q, models, centers)
        listOfPairs = {i: {j: list(pl[i][j]) for j in range(i + 1, k)} for i in range(k - 1)}
        universe = [set(range(n_units)) for _ in range(k)]
        out[name] = (Q, listOfPairs, universe, centers, motif)
    # a mixed query: the geometric lists of geometric_4 with one list thinned like an interaction constraint would
    Q, lop, uni, cen, motif = out["geometric_4"]
    rng = np.random.Generator(np.random.PCG64(263))
    lop2 = {i: {j: list(v) for j, v in d.items()} for i, d in lop.items()}
    keep = rng.random(len(lop2[1][2])) < 0.4
    lop2[1][2] = [p for p, kp in zip(lop2[1][2], keep) if kp]
    Q2 = dict(Q)
    Q2["type"] = "mixed"
    out["mixed_4"] = (Q2, lop2, uni, cen, motif)
    # symbolic queries on random pair lists, with "full" lists and universes that do not contain every unit
    def rand_pairs(n):
        a = rng.integers(0, 120, n)
        b = rng.integers(0, 120, n)
        return sorted(set((int(x), int(y)) for x, y in zip(a, b) if x != y))
    for name, k, full in (("symbolic_4_full", 4, ((0, 2), (0, 3), (1, 3))), ("symbolic_5", 5, ((0, 4), (1, 3))),
                          ("symbolic_reordered", 4, ((0, 2), (1, 2), (0, 3))), ("symbolic_halt_3", 3, ((0, 2), (1, 2))),
                          ("symbolic_halt_2", 2, ((0, 1),))):
        lop = {i: {} for i in range(k - 1)}
        for i in range(k - 1):
            for j in range(i + 1, k):
                lop[i][j] = "full" if (i, j) in full else rand_pairs(900)
        universe = [set(int(v) for v in rng.choice(120, 100, replace=False)) for _ in range(k)]
        Q = {"type": "symbolic", "numpositions": k}
        out[name] = (Q, lop, universe, None, None)
    return out


def main_possibilities():
    rs = import_reference_search()
    out = {}
    for name, (Q, listOfPairs, universe, centers, motif) in possibility_scenarios().items():
        k = Q["numpositions"]
        case = {"type": Q["type"], "numpositions": k, "motif": motif,
                "listOfPairs": {"%d,%d" % (i, j): (v if v == "full" else [[int(a), int(b)] for a, b in v])
                                for i, d in listOfPairs.items() for j, v in d.items()},
                "universe": [sorted(int(v) for v in u) for u in universe]}
        if centers is not None:
            case["centers"] = centers.tolist()
            case["distance"] = Q["distance"].tolist()
            case["cutoff"] = Q["cutoff"]
        ifedata = {"units": [{"centers": c} for c in centers]} if centers is not None else {"units": []}
        Q.update({"MAXTIME": 1e9, "CPUTimeUsed": 0, "errorMessage": [], "halt": False})
        with contextlib.redirect_stdout(io.StringIO()):
            if k > 2 or listOfPairs[0][1] != "full":
                lop, uni, perm = rs.reorderPositions(listOfPairs, universe)
            else:
                lop, uni, perm = listOfPairs, universe, list(range(k))
            if Q["type"] != "symbolic":
                Q["permutedDistance"] = np.zeros((k, k))
                for i in range(k):
                    for j in range(k):
                        Q["permutedDistance"][i][j] = Q["distance"][perm[i]][perm[j]]
            sel = rs.buildSecondElementList(Q, lop, uni, ifedata)
            Q, poss = rs.getPossibilities(Q, ifedata, perm, sel, k, lop)
        case["perm"] = [int(p) for p in perm]
        case["possibilities_sorted"] = sorted([int(v) for v in p] for p in poss)
        case["halt"] = bool(Q["halt"])
        case["errorMessage"] = list(Q["errorMessage"][:1])
        print(name, "perm", perm, "possibilities", len(poss), "halt", Q["halt"])
        out[name] = case
    dump("possibilities.json.gz", out)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "possibilities":
    main_possibilities()
