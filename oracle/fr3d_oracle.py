"""CPU ORACLE — test infrastructure, NOT product code.

A plain Python/numpy restatement of fr3d-python's pairwise nucleotide
annotator and geometric-discrepancy engine (the hot path named in
BASELINE.json).  Only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs may import this module; the product package
(fr3d_python_b200) never does and has no CPU fallback.

Parity status: PINNED.  tests/test_oracle_golden.py checks every function here
against tests/golden/*.json.gz, which tools/make_golden.py produced by running
the UNMODIFIED reference (imported from /root/reference in the dev container)
on structures built from the reference's own test data, and against the known
answers the reference tests hold (tests/discrepancy_tests.py:301-335,
tests/superpositions_test.py:11-119, the Matlab golden frame files).

Every function cites the reference lines it restates
(NA = fr3d/classifiers/NA_pairwise_interactions.py).
Input is the plain "structure spec" (see tools/ref_fixtures.py); this module
reads the exported DATA tables under fr3d_python_b200/tables/ but imports no
product code.
"""

import gzip
import json
import math
import os
from collections import defaultdict

import numpy as np

_TABLES = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       "fr3d_python_b200", "tables")


def _load(name):
    path = os.path.join(_TABLES, name)
    if name.endswith(".gz"):
        with gzip.open(path, "rt") as f:
            return json.load(f)
    with open(path) as f:
        return json.load(f)


DEFS = _load("definitions.json")
PLANES = _load("planes.json")
HBONDS = _load("hbonds.json")
_CUT_RAW = _load("cutoffs.json")

NAbaseheavyatoms = DEFS["NAbaseheavyatoms"]
NAbasehydrogens = DEFS["NAbasehydrogens"]
NAbasecoordinates = DEFS["NAbasecoordinates"]
NAbaseMassiveAndHydrogens = DEFS["NAbaseMassiveAndHydrogens"]
NAbaseatoms = {s: set(NAbaseheavyatoms[s] + NAbasehydrogens[s]) for s in NAbaseheavyatoms}
MODIFIED = _load("modified.json.gz")     # fr3d/data/mapping.py:26-83 applied to atom_mappings.txt


def nt_nt_cutoffs():
    """class_limits_2023.nt_nt_cutoffs as nested dicts in the original order."""
    out = {}
    for combo, ilist in _CUT_RAW:
        out[combo] = {}
        for inter, slist in ilist:
            out[combo][inter] = {}
            for sub, cut in slist:
                out[combo][inter][sub] = cut
    return out


near_discrepancy_cutoff = 1.0     # NA:71
nt_nt_screen_distance = 12        # NA:69


# ----------------------------------------------------------------- geometry --

def besttransformation(set1, set2):
    """fr3d/geometry/superpositions.py:10-89."""
    set1 = np.asarray(set1, dtype=float)
    set2 = np.asarray(set2, dtype=float)
    assert len(set1) == len(set2), 'Lengths must match'
    length = len(set1)
    assert length > 0, 'Must not give empty matrices'
    mean1 = np.sum(set1, axis=0) / float(length)
    mean2 = np.sum(set2, axis=0) / float(length)
    dev1 = set1 - mean1
    dev2 = set2 - mean2
    A = np.dot(dev2.T, dev1)
    V, diagS, Wt = np.linalg.svd(A)
    I = np.identity(3)
    d = np.linalg.det(np.dot(Wt.T, V.T))
    if np.isclose(d, -1.0):
        I[2, 2] = d
    U = np.dot(np.dot(Wt.T, I), V.T)
    new1 = np.dot(dev1, U)
    sse = float(np.sum(np.power(new1 - dev2, 2)))          # RMSD.py:27
    rmsd = float(np.sqrt(np.sum(np.power(new1 - dev2, 2)) / length))  # RMSD.py:15
    return U, new1, mean1, rmsd, sse, mean2


def besttransformation_weighted(set1, set2, weights=(1.0,)):
    """fr3d/geometry/superpositions.py:92-121 (weights only shape A; sse is unweighted)."""
    set1 = np.asarray(set1, dtype=float)
    set2 = np.asarray(set2, dtype=float)
    assert len(set1) == len(set2)
    length = len(set1)
    assert length > 0
    if len(weights) == len(set1):
        diagonal = np.diag(weights)
    else:
        diagonal = np.diag(np.ones(len(set1)))
    mean1 = np.sum(set1, axis=0) / float(length)
    mean2 = np.sum(set2, axis=0) / float(length)
    dev1 = set1 - mean1
    dev2 = set2 - mean2
    A = np.dot(dev2.T, np.dot(diagonal, dev1))
    V, diagS, Wt = np.linalg.svd(A)
    I = np.identity(3)
    d = np.linalg.det(np.dot(Wt.T, V.T))
    if np.isclose(d, -1.0):
        I[2, 2] = d
    U = np.dot(np.dot(Wt.T, I), V.T)
    new1 = np.dot(dev1, U)
    sse = float(np.sum(np.power(new1 - dev2, 2)))
    rmsd = float(np.sqrt(np.sum(np.power(new1 - dev2, 2)) / length))
    return U, new1, mean1, rmsd, sse


def angle_of_rotation(rotation_matrix):
    """fr3d/geometry/angleofrotation.py:4-7."""
    value = (np.trace(rotation_matrix) - 1.0) / 2.0
    value = np.clip(value, -1, 1)
    return float(np.arccos(value))


def _has_rot(r):
    return r is not None and np.asarray(r).shape[0] > 0


def matrix_discrepancy(centers1, rotations1, centers2, rotations2, angle_weight=None, center_weight=None):
    """fr3d/geometry/discrepancy.py:165-233."""
    n = len(centers1)
    assert len(centers2) == n
    assert len(rotations1) == n
    assert len(rotations2) == n
    assert n >= 2
    if not angle_weight:
        angle_weight = 1.0
    if not center_weight:
        center_weight = [1.0] * n
    if n > 2:
        U, new1, mean1, rmsd, sse = besttransformation_weighted(centers1, centers2, center_weight)
        orientation_error = 0
        for r1, r2 in zip(rotations1, rotations2):
            if _has_rot(r1) and _has_rot(r2):
                angle = angle_of_rotation(np.dot(np.dot(U, r2), np.transpose(r1)))
                orientation_error += np.square(angle)
        return float(np.sqrt(sse + angle_weight * orientation_error) / n)
    return _discrepancy_n2(centers1, rotations1, centers2, rotations2, angle_weight)


def _discrepancy_n2(centers1, rotations1, centers2, rotations2, angle_weight):
    """fr3d/geometry/discrepancy.py:206-231 (n == 2 closed form)."""
    c1 = [np.asarray(c, dtype=float) for c in centers1]
    c2 = [np.asarray(c, dtype=float) for c in centers2]
    r1 = [np.asarray(r, dtype=float) for r in rotations1]
    r2 = [np.asarray(r, dtype=float) for r in rotations2]
    R1 = np.dot(r1[1].T, r1[0])
    R2 = np.dot(r2[0].T, r2[1])
    ang1 = angle_of_rotation(np.dot(R1, R2))
    ang2 = angle_of_rotation(np.dot(R1.T, R2.T))
    T1 = np.dot(c1[1] - c1[0], r1[0])
    T2 = np.dot(c1[0] - c1[1], r1[1])
    S1 = np.dot(c2[1] - c2[0], r2[0])
    S2 = np.dot(c2[0] - c2[1], r2[1])
    D1 = T1 - S1
    D2 = T2 - S2
    d = np.sqrt(D1[0] ** 2 + D1[1] ** 2 + D1[2] ** 2 + (angle_weight * ang1) ** 2)
    d += np.sqrt(D2[0] ** 2 + D2[1] ** 2 + D2[2] ** 2 + (angle_weight * ang2) ** 2)
    return float(d * 0.17677669529663687)


def matrix_discrepancy_cutoff(centers1, rotations1, centers2, rotations2, cutoff,
                              angle_weight=None, center_weight=None):
    """fr3d/geometry/discrepancy.py:235-313 (None once the running error passes the cutoff)."""
    n = len(centers1)
    assert len(centers2) == n
    assert len(rotations1) == n
    assert len(rotations2) == n
    assert n >= 2
    if not angle_weight:
        angle_weight = [1] * n
    if not center_weight:
        center_weight = [1] * n
    if n > 2:
        temp_cutoff = (n * cutoff) ** 2
        U, new1, mean1, rmsd, sse = besttransformation_weighted(centers1, centers2, center_weight)
        total_error = sse
        if total_error > temp_cutoff:
            return None
        for r1, r2 in zip(rotations1, rotations2):
            if _has_rot(r1) and _has_rot(r2):
                angle = angle_of_rotation(np.dot(np.dot(U, r2), np.transpose(r1)))
                total_error += np.square(angle)
                if total_error > temp_cutoff:
                    return None
        return float(np.sqrt(total_error) / n)
    return _discrepancy_n2(centers1, rotations1, centers2, rotations2, angle_weight[0])


def all_vs_all_discrepancy(centers, rotations, has_rot=None):
    """The double loop of fr3d/search/FR3D.py:433-442 (diagonal left at 0)."""
    N = len(centers)
    D = np.zeros((N, N))
    for i in range(N):
        for j in range(i + 1, N):
            ri = [rotations[i][k] if has_rot is None or has_rot[i][k] else np.empty((0, 0))
                  for k in range(len(centers[i]))]
            rj = [rotations[j][k] if has_rot is None or has_rot[j][k] else np.empty((0, 0))
                  for k in range(len(centers[j]))]
            D[i, j] = matrix_discrepancy(list(centers[i]), ri, list(centers[j]), rj)
            D[j, i] = D[i, j]
    return D


# ------------------------------------------------------------ nucleotides ----

class Nt(object):
    """What the annotator reads from a fr3d Component (fr3d/data/components.py:118-202)."""

    def __init__(self, pdb, d):
        self.pdb = pdb
        self.model = d["model"]
        self.chain = d["chain"]
        self.sequence = d["sequence"]
        self.number = d["number"]
        self.index = d["index"]
        self.symmetry = d["symmetry"]
        self.alt_id = d.get("alt_id")
        self.insertion_code = d.get("insertion_code")
        self.atom_names = [a[0] for a in d["atoms"]]
        self.atom_xyz = [np.array(a[1:4], dtype=float) for a in d["atoms"]]
        self.rotation = None
        self.base_center = None
        self.fitted = None
        self._fit_frame()
        self._infer_hydrogens()

    def unit_id(self):
        """fr3d/unit_ids.py:31-64 with the Component fields (components.py:1030-1045)."""
        fields = [self.pdb, self.model, self.chain, self.sequence, self.number, None,
                  self.alt_id, self.insertion_code, self.symmetry]
        ordered = ['' if v is None else str(v) for v in fields]
        defaults = {'symmetry': '1_555'}
        possible = ['symmetry', 'insertion_code', 'alt_id', 'atom_name', 'component_number',
                    'component_id', 'chain', 'model']
        while (possible and ordered[-1] == defaults.get(possible[0])) or not ordered[-1]:
            ordered.pop()
            possible.pop(0)
        return '|'.join(ordered)

    def center(self, name):
        """AtomProxy lookup (fr3d/data/base.py:150-166): the atom of that name, the mean
        if several, empty array if none."""
        if name == "base":
            return self.base_center if self.base_center is not None else np.array([])
        coords = [x for n, x in zip(self.atom_names, self.atom_xyz) if n == name]
        if not coords:
            return np.array([])
        if len(coords) == 1:
            return coords[0]
        return np.average(coords, axis=0)

    def mapped(self, name):
        """get_one_atom_coordinates / get_atom_coordinates (NA:324-371): parent atom name -> the name
        this (possibly modified) residue uses."""
        m = MODIFIED.get(self.sequence)
        if m is not None:
            return m["parent_atom_to_modified"].get(name, name)
        return name

    def base_atom_names(self):
        """get_base_atom_names, NA:1574-1599."""
        if self.sequence in NAbaseheavyatoms:
            return NAbaseatoms[self.sequence]
        m = MODIFIED.get(self.sequence)
        if m is not None:
            p2m = m["parent_atom_to_modified"]
            return set(p2m[a] for a in NAbaseatoms[m["parent"]] if a in p2m)
        return set()

    def _fit_frame(self):
        """Component.calculate_rotation_matrix, components.py:273-349."""
        R = []
        S = []
        if self.sequence in NAbaseheavyatoms:
            heavy = NAbaseheavyatoms[self.sequence]
            std = NAbasecoordinates[self.sequence]
            for n, x in zip(self.atom_names, self.atom_xyz):
                if n in heavy:
                    R.append(x)
                    S.append(std[n])
        elif self.sequence in MODIFIED:
            m = MODIFIED[self.sequence]
            heavy = NAbaseheavyatoms[m["parent"]]
            std = NAbasecoordinates[m["parent"]]
            listed = set(m["base_atoms"])
            for n, x in zip(self.atom_names, self.atom_xyz):
                if n in listed:
                    pa = m["modified_atom_to_parent"][n]
                    if pa in heavy:
                        R.append(x)
                        S.append(std[pa])
        else:
            return
        if len(R) == 0:
            return
        U, fitted, meanR, rmsd, sse, meanS = besttransformation(np.array(R), np.array(S))
        self.rotation = U
        self.fitted = fitted
        if self.sequence in NAbaseheavyatoms:
            self.base_center = meanR                               # components.py:342-344
        else:
            self.base_center = np.subtract(meanR, np.dot(U, meanS))   # components.py:348-349

    def _infer_hydrogens(self):
        """Component.infer_NA_hydrogens, components.py:359-516.  Standard residues: every missing base
        hydrogen at centre + U . h_std.  Modified residues: if NO atom of the residue has an 'H' in its
        name, an ideal-position copy of EVERY mapped base atom (heavy atoms included,
        mapping.py:76-79) is appended (:505-513); otherwise nothing is added."""
        if self.rotation is None:
            return
        if self.sequence in NAbasehydrogens:
            already = set(self.atom_names)
            coords = NAbasecoordinates[self.sequence]
            for h in NAbasehydrogens[self.sequence]:
                if h not in already:
                    new = self.base_center + np.dot(self.rotation, np.array(coords[h]))
                    self.atom_names.append(h)
                    self.atom_xyz.append(np.array(new, dtype=float))
        elif self.sequence in MODIFIED:
            if any('H' in n for n in self.atom_names):
                return
            m = MODIFIED[self.sequence]
            std = NAbasecoordinates[m["parent"]]
            for name in m["base_atoms"]:
                new = self.base_center + np.dot(self.rotation, np.array(std[m["modified_atom_to_parent"][name]]))
                self.atom_names.append(name)
                self.atom_xyz.append(np.array(new, dtype=float))

    def base_points(self):
        """Coordinates of atoms whose name is a base atom, in atom order
        (NA:1811-1817, 2211-2220)."""
        names = self.base_atom_names()
        return [x for n, x in zip(self.atom_names, self.atom_xyz) if n in names]


def build_nts(spec):
    return [Nt(spec["pdb"], d) for d in spec["nts"]]


def get_parent(sequence):
    """NA:1246-1260 (standard residues)."""
    if sequence in ['A', 'C', 'G', 'U']:
        return sequence
    if sequence in ['DA', 'DC', 'DG']:
        return sequence[1]
    if sequence == 'DT':
        return sequence
    if sequence in MODIFIED:
        return MODIFIED[sequence]["parent"]
    return None


def glycosidic(nt):
    """NA:2762-2776."""
    if nt.sequence in ['A', 'G', 'DA', 'DG']:
        return nt.center("N9")
    if nt.sequence in ['C', 'U', 'DC', 'DT']:
        return nt.center("N1")
    if nt.sequence in MODIFIED:
        parent = MODIFIED[nt.sequence]["parent"]
        return nt.center(nt.mapped("N9" if parent in ['A', 'G', 'DA', 'DG'] else "N1"))
    return np.array([])


# Adversarial-test support (tests/adversarial.py): when PROBE is a list, every threshold comparison of the rules
# appends (name, value - threshold), so that a decision boundary found by bisection can be attributed to the
# threshold whose margin changes sign there.  Recording only; no rule reads it.
PROBE = None


def _probe(name, margin):
    if PROBE is not None:
        PROBE.append((name, float(margin)))


def translate_rotate_point(nt, point):
    """NA:1263-1275: (point - centre) . U as a row vector."""
    return list(np.dot(np.subtract(point, nt.base_center), nt.rotation))


# ---------------------------------------------------------------- search ----

HALF_STENCIL = [[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1], [0, 1, -1], [1, 0, 0], [1, 0, 1], [1, 0, -1],
                [1, 1, 0], [1, 1, 1], [1, 1, -1], [1, -1, 0], [1, -1, 1], [1, -1, -1]]   # NA:438


def make_nt_cubes_half(nts, cutoff=nt_nt_screen_distance):
    """NA:410-443, cube lists keyed (x,y,z,model) in first-seen order."""
    cubes = {}
    neighbors = {}
    for k, nt in enumerate(nts):
        c = nt.center("base")
        if len(c) == 3:
            x = math.floor(c[0] / cutoff)
            y = math.floor(c[1] / cutoff)
            z = math.floor(c[2] / cutoff)
            key = (x, y, z, nt.model)
            if key in cubes:
                cubes[key].append(k)
            else:
                cubes[key] = [k]
                neighbors[key] = [(x + a, y + b, z + c2, nt.model) for a, b, c2 in HALF_STENCIL]
    return cubes, neighbors


def candidate_pairs(nts, cutoff=nt_nt_screen_distance):
    """The 4-deep loop and screens of NA:661-724; returns oriented (i, j) in visit order."""
    cubes, neighbors = make_nt_cubes_half(nts, cutoff)
    out = []
    for key1 in cubes:
        for key2 in neighbors[key1]:
            if key2 not in cubes:
                continue
            for i in cubes[key1]:
                nt1 = nts[i]
                if len(nt1.center("base")) < 3:
                    continue
                if len(glycosidic(nt1)) < 3:
                    continue
                for j in cubes[key2]:
                    nt2 = nts[j]
                    if key1 == key2:
                        if nt1.chain > nt2.chain:
                            continue
                        elif nt1.chain == nt2.chain and nt1.index > nt2.index:
                            continue
                    if len(nt2.center("base")) < 3:
                        continue
                    displacement = abs(nt2.base_center - nt1.base_center)
                    _probe("search |dx|,|dy|,|dz| vs 12", max(displacement) - cutoff)
                    if displacement[0] > cutoff or displacement[1] > cutoff or displacement[2] > cutoff:
                        continue
                    if nt1.number == nt2.number and nt1.sequence == nt2.sequence and nt1.chain == nt2.chain \
                            and nt1.symmetry == nt2.symmetry and nt1.insertion_code == nt2.insertion_code \
                            and nt1.alt_id != nt2.alt_id:
                        continue
                    d = np.linalg.norm(displacement)
                    _probe("search d vs 12", d - cutoff)
                    _probe("search d vs 2", d - 2)
                    if d > cutoff:
                        continue
                    if d < 2:
                        continue
                    out.append((i, j))
    return out


# ----------------------------------------------------------------- rules ----

def _inside(lines, x, y, what=None):
    for a, b, c in lines:
        if what is not None:
            _probe(what, a * x + b * y + c)
        if not (a * x + b * y + c > 0):
            return False
    return True


def _ring_parent(parent):
    return {'DA': 'A', 'DC': 'C', 'DG': 'G'}.get(parent, parent)


OXYGENS = ["O2'", "O3'", "O4'", "O5'", "OP1", "OP2"]     # NA:1290


def check_base_oxygen_stack_rings(nt1, nt2, parent1):
    """NA:1278-1473.  Returns (interaction, interaction_reversed)."""
    true_z_cutoff = 3.5
    near_z_cutoff = 3.6
    interaction = ""
    interaction_reversed = ""
    rp = _ring_parent(parent1)
    rings = PLANES["rings"].get(rp)
    ell = PLANES["ellipses"].get(rp)
    oxygen_points = []
    true_found = False
    near_found = False
    zmin = 999
    oxygenmin = None
    for oxygen in OXYGENS:
        p = nt2.center(oxygen)
        if len(p) == 3:
            x, y, z = translate_rotate_point(nt1, p)
            oxygen_points.append([x, y, z, oxygen])
            _probe("sO |z| vs 2", abs(z) - 2)
            _probe("sO |z| vs 3.6", abs(z) - near_z_cutoff)
            _probe("sO |z| vs 3.5", abs(z) - true_z_cutoff)
            if abs(z) < 2:
                continue
            ring5 = False
            ring6 = False
            if abs(z) < near_z_cutoff and rings is not None:
                if rings["divider"] is not None:
                    a, b, c = rings["divider"]
                    _probe("sO ring divider", a * x + b * y + c)
                    if a * x + b * y + c > 0:
                        ring5 = _inside(rings["ring5"], x, y, "sO ring5 line")
                    else:
                        ring6 = _inside(rings["ring6"], x, y, "sO ring6 line")
                else:
                    ring6 = _inside(rings["ring6"], x, y, "sO ring6 line")
            if ring5 or ring6:
                if abs(z) < true_z_cutoff:
                    true_found = True
                else:
                    near_found = True
                if abs(z) < abs(zmin):
                    zmin = z
                    oxygenmin = oxygen
    if true_found:
        if zmin > 0:
            interaction = "s3" + oxygenmin
            interaction_reversed = "s" + oxygenmin + "3"
        else:
            interaction = "s5" + oxygenmin
            interaction_reversed = "s" + oxygenmin + "5"
    elif near_found:
        if zmin > 0:
            interaction = "ns3" + oxygenmin
            interaction_reversed = "ns" + oxygenmin + "3"
        else:
            interaction = "ns5" + oxygenmin
            interaction_reversed = "ns" + oxygenmin + "5"
    else:
        r2min = 999
        for x, y, z, oxygen in oxygen_points:
            near5 = False
            near6 = False
            if abs(z) < true_z_cutoff and ell is not None:
                if rings["divider"] is not None:
                    a, b, c = rings["divider"]
                    if a * x + b * y + c > 0:
                        near5 = _in_ellipse(ell["near5"], x, y)
                    else:
                        near6 = _in_ellipse(ell["near6"], x, y)
                else:
                    near6 = _in_ellipse(ell["near6"], x, y)
            if near5 or near6:
                near_found = True
                r2 = x ** 2 + y ** 2
                if r2 < r2min:
                    r2min = r2
                    zmin = z
                    oxygenmin = oxygen
        if near_found:
            if zmin > 0:
                interaction = "ns3" + oxygenmin
                interaction_reversed = "ns" + oxygenmin + "3"
            else:
                interaction = "ns5" + oxygenmin
                interaction_reversed = "ns" + oxygenmin + "5"
    return interaction, interaction_reversed


def _in_ellipse(e, x, y):
    a, b, cx, cy, r = e
    _probe("sO near-ring ellipse", a * (x - cx) ** 2 + b * (x - cx) * (y - cy) + (y - cy) ** 2 - r)
    return a * (x - cx) ** 2 + b * (x - cx) * (y - cy) + (y - cy) ** 2 < r


def check_convex_hull_atoms(x, y, z, parent):
    """NA:1475-1528."""
    _probe("stacking atom |z| vs 4.5", abs(z) - 4.5)
    if abs(z) < 4.5:
        hull = PLANES["hull"].get(_ring_parent(parent))
        if hull is None:
            return False
        return _inside(hull, x, y, "stacking hull line")
    return False


def return_overlap(atom_names, nt1, nt2, parent):
    """NA:1530-1571: project base atoms of nt2 into the frame of nt1."""
    min_z = 1000
    maxz = -1000
    minz = 1000
    ret = [-100, -100, -100]
    overlap = False
    for atom in atom_names:
        point = nt2.center(atom)
        if len(point) == 3:
            x, y, z = translate_rotate_point(nt1, point)
            inside = check_convex_hull_atoms(x, y, z, parent)
            if abs(z) < abs(min_z):
                min_z = z
                ret = [x, y, z]
            if z < minz:
                minz = z
            if z > maxz:
                maxz = z
            if inside:
                overlap = True
    _probe("stacking atoms on both sides (max z)", maxz)
    _probe("stacking atoms on both sides (min z)", minz)
    if maxz > 0 and minz < 0:
        return False, [-100, -100, -100]
    if overlap:
        return True, ret
    return False, [-100, -100, -100]


def calculate_base_min_distances(nt1, nt2):
    """NA:1799-1829."""
    pts2 = nt2.base_points()
    best = 9999
    for q in nt1.base_points():
        for p in pts2:
            d = np.linalg.norm(np.subtract(p, q))
            if d < best:
                best = d
    return float(best)


def check_base_base_stacking(nt1, nt2, parent1, parent2):
    """NA:1602-1739.  Returns (interaction, interaction_reversed)."""
    true_z_cutoff = 4
    interaction = ""
    interaction_reversed = ""
    reverse = False
    # atom order: the reference iterates a set; any order gives the same result except
    # for exact |z| ties (measure zero).  Sorted for determinism.
    a1 = sorted(nt1.base_atom_names())
    a2 = sorted(nt2.base_atom_names())
    if len(a1) == 0 or len(a2) == 0:
        return "", ""
    nt2on1, coords = return_overlap(a2, nt1, nt2, parent1)
    nt1on2, coords2 = return_overlap(a1, nt2, nt1, parent2)
    if nt2on1 or nt1on2:
        if nt1on2 and not nt2on1:
            normal_Z = np.dot(nt2.rotation.T, nt1.rotation)[2, 2]
            mvd = coords2[2]
            reverse = True
        elif not nt1on2 and nt2on1:
            normal_Z = np.dot(nt1.rotation.T, nt2.rotation)[2, 2]
            mvd = coords[2]
        else:
            normal_Z = np.dot(nt2.rotation.T, nt1.rotation)[2, 2]
            mvd = coords[2]
        if mvd > 0:
            if normal_Z > 0:
                interaction, interaction_reversed = "ns35", "ns53"
            elif normal_Z < 0:
                interaction, interaction_reversed = "ns33", "ns33"
        elif mvd < 0:
            if normal_Z > 0:
                interaction, interaction_reversed = "ns53", "ns35"
            elif normal_Z < 0:
                interaction, interaction_reversed = "ns55", "ns55"
        if reverse:
            interaction, interaction_reversed = interaction_reversed, interaction
        min_distance = calculate_base_min_distances(nt1, nt2)
        _probe("stacking min distance vs 4", abs(min_distance) - true_z_cutoff)
        _probe("stacking min distance vs 1", abs(min_distance) - 1)
        _probe("stacking |normal_Z| vs 0.6", abs(normal_Z) - 0.6)
        _probe("stacking |normal_Z| vs 0.5", abs(normal_Z) - 0.5)
        _probe("stacking normal_Z sign", normal_Z)
        _probe("stacking vertical offset sign", mvd)
        if abs(min_distance) < true_z_cutoff and abs(min_distance) > 1 and abs(normal_Z) > 0.6 \
                and nt2on1 and nt1on2:
            interaction = interaction.replace("n", "")
            interaction_reversed = interaction_reversed.replace("n", "")
        if abs(min_distance) < 1 or abs(normal_Z) < 0.5:
            return "", ""
    return interaction, interaction_reversed


def torsion_angle(p0, p1, p2, p3):
    """NA:2904-2930."""
    b0 = -1.0 * (p1 - p0)
    b1 = p2 - p1
    b2 = p3 - p2
    b1 = b1 / np.linalg.norm(b1)
    v = b0 - np.dot(b0, b1) * b1
    w = b2 - np.dot(b2, b1) * b1
    x = np.dot(v, w)
    y = np.dot(np.cross(b1, v), w)
    return float(np.degrees(np.arctan2(y, x)))


def check_sugar_ribose(nt1, nt2, parent1):
    """NA:2262-2342."""
    o1 = nt1.center(nt1.mapped("O2'"))
    if not len(o1) == 3:
        return ""
    o2 = nt2.center(nt2.mapped("O2'"))
    if not len(o2) == 3:
        return ""
    _probe("SR O2'-O2' vs 3.8", np.linalg.norm(np.subtract(o1, o2)) - 3.8)
    if np.linalg.norm(np.subtract(o1, o2)) > 3.8:
        return ""
    if parent1 in ['A', 'G']:
        base_point = nt1.center(nt1.mapped('N3'))
    elif parent1 in ['C', 'U']:
        base_point = nt1.center(nt1.mapped('O2'))
    else:
        return ""
    if not len(base_point) == 3:
        return ""
    displ = np.subtract(base_point, o2)
    _probe("SR base atom-O2' vs 3.8", np.linalg.norm(displ) - 3.8)
    if np.linalg.norm(displ) > 3.8:
        return ""
    z = abs(float(np.dot(displ, nt1.rotation[:, 2])))
    _probe("SR height vs 2.0", z - 2.0)
    if z > 2.0:
        return ""
    p0, p1 = nt1.center(nt1.mapped("C1'")), nt1.center(nt1.mapped("C2'"))
    p2, p3 = nt2.center(nt2.mapped("C2'")), nt2.center(nt2.mapped("C1'"))
    if len(p0) == 3 and len(p1) == 3 and len(p2) == 3 and len(p3) == 3:
        angle = torsion_angle(p0, p1, p2, p3)
        _probe("SR |torsion| vs 90", abs(angle) - 90)
        return 'cSR' if abs(angle) > 90 else 'tSR'
    return ""


def angle_between_vectors(v1, v2):
    """NA:2873-2881."""
    cosang = np.dot(v1, v2)
    sinang = np.linalg.norm(np.cross(v1, v2))
    return float(180 * np.arctan2(sinang, cosang) / np.pi)


def previous_O3_map(nts):
    """map_unit_id_to_previous_O3, NA:480-525; keyed by nt position.  Missing previous
    O3' -> None (the reference's np.empty([1,3]) placeholder is inert: calculate_hb_angle
    returns None for it, NA:2857-2859)."""
    order = sorted(range(len(nts)),
                   key=lambda k: (nts[k].model, nts[k].symmetry, nts[k].chain, nts[k].index, nts[k].unit_id()))
    out = {}
    prev = None
    for k in order:
        nt = nts[k]
        if prev is not None and nt.model == prev.model and nt.symmetry == prev.symmetry \
                and nt.chain == prev.chain and nt.index == prev.index + 1:
            out[k] = prev.center(prev.mapped("O3'"))
        else:
            out[k] = None
        prev = nt
    return out


def check_base_backbone_interactions(nt1, nt2, previousO3, parent1):
    """NA:1867-2049, including the per-site sticky selection at :1973-2019."""
    phosphate = ""
    ribose = ""
    site_to_oxy = {}
    true_ph = []
    near_ph = []
    true_r = []
    near_r = []
    carbonCutoff, nCarbonCutoff = 4.0, 4.5
    nitrogenCutoff, nNitrogenCutoff = 3.5, 4.0
    angleLimit, nAngleLimit = 130, 110
    ph_ox = [nt2.center(nt2.mapped(n)) for n in ["O5'", 'OP1', 'OP2']]
    r_ox = [nt2.center(nt2.mapped(n)) for n in ["O2'", "O4'"]]
    if previousO3 is not None and len(previousO3) == 3 and previousO3.any():
        ph_ox.append(previousO3)
    P = nt2.center(nt2.mapped("P"))
    if P.any():
        _probe("BPh P off-plane vs 4.5", abs(translate_rotate_point(nt1, P)[2]) - 4.5)
        if abs(translate_rotate_point(nt1, P)[2]) > 4.5:
            ph_ox = []
    cutoff = nCutoff = None
    for site in NAbaseMassiveAndHydrogens[parent1]:
        H = nt1.center(nt1.mapped(site[0]))
        M = nt1.center(nt1.mapped(site[1]))
        if "C" in site[1]:
            cutoff, nCutoff = carbonCutoff, nCarbonCutoff
        elif "N" in site[1]:
            cutoff, nCutoff = nitrogenCutoff, nNitrogenCutoff
        for i, ox in enumerate(ph_ox):
            if ox.any() and len(M) == 3 and len(H) == 3 and len(ox) == 3:
                ang = angle_between_vectors(np.subtract(M, H), np.subtract(ox, H))
                if ang:
                    dist = float(np.linalg.norm(np.subtract(M, ox)))
                    if dist:
                        _probe("BPh angle vs 130", ang - angleLimit)
                        _probe("BPh angle vs 110", ang - nAngleLimit)
                        _probe("BPh distance vs 4.0 (C) / 3.5 (N)", dist - cutoff)
                        _probe("BPh distance vs 4.5 (C) / 4.0 (N)", dist - nCutoff)
                        if ang > angleLimit and dist < cutoff:
                            q = (cutoff - dist) + (ang - angleLimit) / 20.0
                            true_ph.append((-q, site[2], i))
                            site_to_oxy.setdefault(site[2], set()).add(i)
                        elif ang > nAngleLimit and dist < nCutoff:
                            q = (nCutoff - dist) + (ang - nAngleLimit) / 20.0
                            near_ph.append((-q, "n" + site[2]))
        for ox in r_ox:
            if ox.any() and len(M) == 3 and len(H) == 3 and len(ox) == 3:
                ang = angle_between_vectors(np.subtract(M, H), np.subtract(ox, H))
                if ang:
                    dist = float(np.linalg.norm(np.subtract(M, ox)))
                    if dist:
                        _probe("BR angle vs 130", ang - angleLimit)
                        _probe("BR angle vs 110", ang - nAngleLimit)
                        _probe("BR distance vs 4.0 (C) / 3.5 (N)", dist - cutoff)
                        _probe("BR distance vs 4.5 (C) / 4.0 (N)", dist - nCutoff)
                        if ang > angleLimit and dist < cutoff:
                            q = (cutoff - dist) + (ang - angleLimit) / 20.0
                            true_r.append((-q, site[3]))
                        elif ang > nAngleLimit and dist < nCutoff:
                            q = (nCutoff - dist) + (ang - nAngleLimit) / 20.0
                            near_r.append((-q, "n" + site[3]))
        if len(true_ph) == 1:
            phosphate = true_ph[0][1]
        elif len(site_to_oxy) > 1:
            if '7BPh' in site_to_oxy and '9BPh' in site_to_oxy:
                if len(site_to_oxy['7BPh'] | site_to_oxy['9BPh']) > 1:
                    phosphate = "8BPh"
            elif '3BPh' in site_to_oxy and '5BPh' in site_to_oxy:
                if len(site_to_oxy['3BPh'] | site_to_oxy['5BPh']) > 1:
                    phosphate = "4BPh"
            if not phosphate:
                phosphate = sorted(true_ph)[0][1]
        elif len(true_ph) > 1:
            phosphate = sorted(true_ph)[0][1]
        elif len(near_ph) == 1:
            phosphate = near_ph[0][1]
        elif len(near_ph) > 1:
            phosphate = sorted(near_ph)[0][1]
        if len(true_r) == 1:
            ribose = true_r[0][1]
        elif len(true_r) > 1:
            ribose = sorted(true_r)[0][1]
        elif len(near_r) == 1:
            ribose = near_r[0][1]
        elif len(near_r) > 1:
            ribose = sorted(near_r)[0][1]
    return phosphate, ribose


def calculate_basepair_gap(nt1, nt2):
    """NA:2189-2259: min |z| (frame of nt1) over the 3 base atoms of nt2 nearest centre 1."""
    displacements = []
    distances = []
    for p in nt2.base_points():
        v = np.subtract(p, nt1.base_center)
        displacements.append(v)
        distances.append(np.linalg.norm(v))
    indices = np.argsort(distances)
    m = min(3, len(indices))
    gap12 = 100
    for k in range(m):
        z = abs(float(np.dot(displacements[indices[k]], nt1.rotation[:, 2])))
        if z < gap12:
            gap12 = z
    return gap12


def check_coplanar(nt1, nt2, pair_data):
    """NA:2051-2186."""
    pair_data["coplanar"] = False
    pair_data["coplanar_value"] = None
    gap12 = calculate_basepair_gap(nt1, nt2)
    pair_data["gap12"] = gap12
    _probe("coplanar gap12 vs 1.5179", gap12 - 1.5179)
    if gap12 >= 1.5179:
        return pair_data
    min_distance = calculate_base_min_distances(nt1, nt2)
    pair_data["min_distance"] = min_distance
    _probe("coplanar min distance vs 3.4589", min_distance - 3.4589)
    _probe("coplanar min distance vs 2.4589", min_distance - 2.4589)
    if min_distance >= 3.4589:
        return pair_data
    if min_distance >= 2.4589 and nt1.sequence in ['A', 'C', 'G', 'U'] and nt2.sequence in ['A', 'C', 'G', 'U']:
        return pair_data
    cd = np.subtract(nt1.base_center, nt2.base_center)
    cd = cd / np.linalg.norm(cd)
    dot1 = abs(float(np.dot(cd, nt1.rotation[:, 2])))
    _probe("coplanar |c.n1| vs 0.3381", dot1 - 0.3381)
    if dot1 >= 0.3381:
        return pair_data
    dot2 = abs(float(np.dot(cd, nt2.rotation[:, 2])))
    _probe("coplanar |c.n2| vs 0.3381", dot2 - 0.3381)
    if dot2 >= 0.3381:
        return pair_data
    dot3 = abs(float(np.dot(nt1.rotation[:, 2], nt2.rotation[:, 2])))
    _probe("coplanar |n1.n2| vs 0.7757", dot3 - 0.7757)
    if dot3 <= 0.7757:
        return pair_data
    gap21 = calculate_basepair_gap(nt2, nt1)
    _probe("coplanar gap21 vs 1.5179 (value only)", gap21 - 1.5179)
    # note: the reference does NOT store gap21 into pair_data here (NA:2115), so
    # check_basepair_cutoffs recomputes it (same value).

    def pw(v, a, b, c, s1, s2):
        if v < a:
            return 1
        if v < b:
            return 1 + (v - a) * s1
        if v < c:
            return 0.5 + (v - b) * s2
        return 0

    vals = [pw(gap12, 0.5062, 0.9775, 1.5179, -1.0609, -0.9252),
            pw(gap21, 0.5062, 0.9775, 1.5179, -1.0609, -0.9252),
            pw(dot1, 0.1139, 0.2193, 0.3381, -4.7408, -4.2103),
            pw(dot2, 0.1139, 0.2193, 0.3381, -4.7408, -4.2103),
            pw(-dot3, -0.9509, -0.8835, -0.7757, -7.4217, -4.6390),
            pw(min_distance, 1.8982, 2.1357, 2.4859, -2.1050, -1.4280)]
    pair_data["coplanar"] = True
    pair_data["coplanar_value"] = min(vals)
    return pair_data


def focus_basepair_cutoffs(basepair_cutoffs, interactions):
    """NA:88-136."""
    focused = {}
    lower = set()
    if interactions:
        for i in interactions:
            lower.add(i.lower())
    else:
        for combo in basepair_cutoffs:
            for i in basepair_cutoffs[combo]:
                lower.add(i.lower())
    for combo in basepair_cutoffs:
        focused[combo] = {1: {}, -1: {}}
        for inter in basepair_cutoffs[combo]:
            family = inter.lower()[0:3]
            if family in lower:
                subcat = list(basepair_cutoffs[combo][inter].keys())[0]
                sgn = 1 if basepair_cutoffs[combo][inter][subcat]["normalmin"] > 0 else -1
                focused[combo][sgn][inter] = dict(basepair_cutoffs[combo][inter])
    return focused


def store_basepair_quality(cutoff_distance, pair_data, LW, hydrogen_bonds):
    """NA:2345-2361."""
    q = {'cutoff_distance': cutoff_distance, 'max_gap': max(pair_data["gap12"], pair_data["gap21"]),
         'atoms1': [], 'atoms2': []}
    LW_clean = LW.replace("n", "")
    if LW_clean in hydrogen_bonds:
        for donor, hydrogen, acceptor, direction in hydrogen_bonds[LW_clean]:
            if direction == "12":
                q['atoms1'].append(hydrogen)
                q['atoms2'].append(acceptor)
            else:
                q['atoms2'].append(hydrogen)
                q['atoms1'].append(acceptor)
    return q


def check_basepair_cutoffs(nt1, nt2, pair_data, cutoffs, hydrogen_bonds):
    """NA:2364-2759 with datapoint = None.  Returns (LW, subcategory, quality)."""
    displ = pair_data["displ12"]
    quality = {}
    _probe("basepair |z| vs 3.6", abs(displ[2]) - 3.6)
    if abs(displ[2]) > 3.6:
        return "", "", quality
    M = np.dot(nt1.rotation.T, nt2.rotation)
    normal_Z = M[2, 2]
    _probe("basepair normal_Z sign", normal_Z)
    sgn = 1 if normal_Z > 0 else -1
    cmax = near_discrepancy_cutoff
    ok = []
    for inter in cutoffs[sgn]:
        for sub in cutoffs[sgn][inter]:
            cut = cutoffs[sgn][inter][sub]
            cd = 0.0
            cd += max(0, cut['xmin'] - displ[0])
            cd += max(0, displ[0] - cut['xmax'])
            if cd >= cmax:
                continue
            cd += max(0, cut['ymin'] - displ[1])
            cd += max(0, displ[1] - cut['ymax'])
            if cd >= cmax:
                continue
            if 'radiusmax' in cut:
                radius = math.sqrt(displ[0] ** 2 + displ[1] ** 2)
                cd += max(0, radius - cut['radiusmax'])
            if cd >= cmax:
                continue
            cd += max(0, cut['zmin'] - displ[2])
            cd += max(0, displ[2] - cut['zmax'])
            if cd >= cmax:
                continue
            cd += 3 * max(0, cut['normalmin'] - normal_Z)
            cd += 3 * max(0, normal_Z - cut['normalmax'])
            _probe("basepair box, cutoff_distance vs 1.0 (before angle)", cd - cmax)
            if PROBE is not None:
                for nm, v in (("x", displ[0]), ("y", displ[1]), ("z", displ[2]), ("normal", normal_Z)):
                    _probe("basepair box edge %smin" % nm, v - cut[nm + 'min'])
                    _probe("basepair box edge %smax" % nm, v - cut[nm + 'max'])
            if cd < cmax:
                ok.append((inter, sub, cut, cd))
    if len(ok) == 0:
        return "", 9999, quality
    angle = math.atan2(M[1, 1], M[0, 1]) * 57.29577951308232 - 90
    if angle <= -90:
        angle += 360
    ok2 = []
    for inter, sub, cut, cd in ok:
        if cut['anglemin'] < cut['anglemax']:
            cd += 0.05 * max(0, cut['anglemin'] - angle)
            cd += 0.05 * max(0, angle - cut['anglemax'])
        else:
            cd += 0.05 * min(max(0, cut["anglemin"] - angle), max(0, angle - cut["anglemax"]))
        _probe("basepair box, cutoff_distance vs 1.0 (with angle)", cd - cmax)
        _probe("basepair box angle min", angle - cut['anglemin'])
        _probe("basepair box angle max", angle - cut['anglemax'])
        if cd < cmax:
            ok2.append((inter, sub, cut, cd))
    if len(ok2) == 0:
        return "", 9999, quality
    if 'gap12' not in pair_data:
        pair_data["gap12"] = calculate_basepair_gap(nt1, nt2)
    if 'gap21' not in pair_data:
        pair_data["gap21"] = calculate_basepair_gap(nt2, nt1)
    if 'min_distance' not in pair_data:
        pair_data["min_distance"] = calculate_base_min_distances(nt1, nt2)
    match = []
    near_match = []
    for inter, sub, cut, cd in ok2:
        if cut['gapmax'] > 0.1:
            cd += 4 * max(0, max(pair_data["gap12"], pair_data["gap21"]) - cut['gapmax'])
            _probe("basepair gap vs gapmax", max(pair_data["gap12"], pair_data["gap21"]) - cut['gapmax'])
        _probe("basepair cutoff_distance vs 1.0 (with gap)", cd - near_discrepancy_cutoff)
        _probe("basepair min distance vs 4.1", pair_data['min_distance'] - 4.1)
        _probe("basepair min distance vs 3.75", pair_data['min_distance'] - 3.75)
        one_hbond = False
        if inter == 'cSs' and pair_data['parent2'] in ['C', 'U']:
            one_hbond = True
        elif inter == 'csS' and pair_data['parent2'] == 'U':
            one_hbond = True
        if cd > 0:
            if cd < near_discrepancy_cutoff and not inter.startswith("n"):
                if pair_data['min_distance'] < 4.1 or one_hbond:
                    near_match.append((inter, sub, cd))
        elif inter.startswith("n"):
            near_match.append((inter, sub, cd))
        elif pair_data['min_distance'] > 3.75 and not one_hbond:
            near_match.append((inter, sub, cd))
        else:
            match.append((inter, sub, cd))
    if len(match) > 0:
        match = sorted(match, key=lambda x: (x[1], len(x[0])))
    elif len(near_match) > 0:
        near_match = sorted(near_match, key=lambda x: x[2])
        match = [near_match[0]]
    else:
        return "", 9999, quality
    # with datapoint None, LW_bond_rank is empty, so every branch of NA:2691-2759 uses match[0]
    inter, sub, cd = match[0]
    LW = "n" + inter if (cd > 0 and "n" not in inter) else inter
    quality = store_basepair_quality(cd, pair_data, LW, hydrogen_bonds)
    return LW, sub, quality


def reverse_edges(inter):
    """NA:446-465."""
    if len(inter) <= 2:
        return inter
    if inter == 'N/A':
        return inter
    if len(inter) == 3:
        return inter[0] + inter[2] + inter[1]
    if len(inter) == 4 and inter[0] in 'n!':
        return inter[0] + inter[1] + inter[3] + inter[2]
    if len(inter) == 4:
        return inter[0] + inter[2] + inter[1] + inter[3]
    if len(inter) == 5:
        return inter[0] + inter[1] + inter[3] + inter[2] + inter[4]
    return inter[0:(len(inter) - 2)] + inter[len(inter) - 1] + inter[len(inter) - 2]


def check_for_two_interactions_on_same_edge(unit_id_to_basepairs):
    """NA:528-632."""
    remove_pairs = set()
    make_near_pairs = set()
    for unit_id, basepairs in unit_id_to_basepairs.items():
        if len(basepairs) > 1:
            for i in range(len(basepairs) - 1):
                interaction_1, quality_1, unit_id_1 = basepairs[i]
                e1 = interaction_1.replace("n", "")[1].lower()
                if (unit_id, unit_id_1) in remove_pairs:
                    continue
                for j in range(i + 1, len(basepairs)):
                    interaction_2, quality_2, unit_id_2 = basepairs[j]
                    if (unit_id, unit_id_2) in remove_pairs:
                        continue
                    if (unit_id, unit_id_2) in make_near_pairs:
                        continue
                    e2 = interaction_2.replace("n", "")[1].lower()
                    if not e1 == e2:
                        continue
                    common = set(quality_1['atoms1']) & set(quality_2['atoms1'])
                    if len(common) > 0:
                        names = "".join(common)
                        if "H" in names or len(common) > 1:
                            n1 = interaction_1.startswith("n")
                            n2 = interaction_2.startswith("n")
                            if n1 and n2:
                                if quality_1['cutoff_distance'] < quality_2['cutoff_distance']:
                                    if quality_2['cutoff_distance'] > 0.5 * near_discrepancy_cutoff:
                                        remove_pairs.add((unit_id, unit_id_2))
                                        remove_pairs.add((unit_id_2, unit_id))
                                else:
                                    if quality_1['cutoff_distance'] > 0.5 * near_discrepancy_cutoff:
                                        remove_pairs.add((unit_id, unit_id_1))
                                        remove_pairs.add((unit_id_1, unit_id))
                            elif not n1 and not n2:
                                if quality_1['max_gap'] < quality_2['max_gap']:
                                    make_near_pairs.add((unit_id, unit_id_2))
                                    make_near_pairs.add((unit_id_2, unit_id))
                                else:
                                    make_near_pairs.add((unit_id, unit_id_1))
                                    make_near_pairs.add((unit_id_1, unit_id))
                            else:
                                if quality_2['cutoff_distance'] > 0.5 * near_discrepancy_cutoff:
                                    remove_pairs.add((unit_id, unit_id_2))
                                    remove_pairs.add((unit_id_2, unit_id))
                                elif quality_1['cutoff_distance'] > 0.5 * near_discrepancy_cutoff:
                                    remove_pairs.add((unit_id, unit_id_1))
                                    remove_pairs.add((unit_id_1, unit_id))
    return remove_pairs, make_near_pairs


def calculate_crossing_numbers(nts, interaction_to_pair_list):
    """NA:1057-1179; pairs are nt positions here, chains keyed by chain id only (Q7)."""
    chain_max = defaultdict(lambda: 0)
    for nt in nts:
        chain_max[nt.chain] = max(chain_max[nt.chain], nt.index)
    chain_cww = defaultdict(list)
    for inter in ['cWW', 'cWw', 'cwW', 'acWW', 'acWw', 'acwW']:
        for a, b in interaction_to_pair_list.get(inter, []):
            if nts[a].chain == nts[b].chain:
                p1 = get_parent(nts[a].sequence)
                p2 = get_parent(nts[b].sequence)
                if (p1 or "") + (p2 or "") in ['AU', 'UA', 'CG', 'GC', 'GU', 'UG']:
                    i1, i2 = nts[a].index, nts[b].index
                    chain_cww[nts[a].chain].append((i1, i2) if i1 < i2 else (i2, i1))
    nested = {}
    for chain in chain_cww:
        pairs = sorted(chain_cww[chain], key=lambda p: (p[1] - p[0], p[0]))
        ends = list(range(0, chain_max[chain] + 1))
        for i1, i2 in pairs:
            i = i1 + 1
            while i < i2 and ends[i] > i1 and ends[i] < i2:
                i += 1
            if i == i2:
                ends[i1] = i2
                ends[i2] = i1
        nested[chain] = ends
    out = defaultdict(list)
    for inter in list(interaction_to_pair_list.keys()):
        if inter == "":
            continue
        for a, b in interaction_to_pair_list[inter]:
            crossing = 0
            if nts[a].chain == nts[b].chain and nts[a].chain in nested:
                i1, i2 = sorted([nts[a].index, nts[b].index])
                for i in range(i1 + 1, i2):
                    j = nested[nts[a].chain][i]
                    if j < i1 or j > i2:
                        crossing += 1
            out[inter].append((a, b, crossing))
            if inter in ["s33", "s35", "s53", "s55", "cp", "ns33", "ns35", "ns53", "ns55"]:
                out[reverse_edges(inter)].append((b, a, crossing))
            elif inter[0] in ["c", "t"]:
                out[reverse_edges(inter)].append((b, a, crossing))
            elif inter[0:2] in ["nc", "nt"]:
                out[reverse_edges(inter)].append((b, a, crossing))
    return out


def annotate_covalent_connections(nts, out, category_to_interactions):
    """NA:1181-1210."""
    groups = defaultdict(list)
    for k, nt in enumerate(nts):
        groups[(nt.model + " " + nt.symmetry + " " + nt.chain, nt.index)].append(k)
    keys = sorted(groups.keys())
    for i in range(len(keys) - 1):
        k1, k2 = keys[i], keys[i + 1]
        if k1[0] == k2[0]:
            dist = k2[1] - k1[1]
            for a in groups[k1]:
                for b in groups[k2]:
                    inter = "p_" + str(dist)
                    out[inter].append((a, b, None))
                    category_to_interactions["covalent"].add(inter)


BASEPAIR_COMBOS = set(['A,A', 'A,C', 'A,G', 'A,U', 'C,C', 'G,C', 'C,U', 'G,G', 'G,U', 'U,U',
                       'A,DT', 'C,DT', 'G,DT', 'DT,DT'])     # NA:654


def annotate(spec, categories, nts=None, focused=None):
    """annotate_nt_nt_in_structure (NA:1213-1243) + annotate_nt_nt_interactions (NA:635-1055).
    focused: the focused_basepair_cutoffs argument of the reference (default: focused from the shipped tables).

    Returns (interaction_to_list_of_tuples, category_to_interactions, traces) where the
    tuples hold nt positions (index into spec["nts"]) instead of unit-id strings and
    traces holds per-rule results in call order (same layout as tools/make_golden.py).
    """
    if nts is None:
        nts = build_nts(spec)
    if focused is None:
        focused = focus_basepair_cutoffs(nt_nt_cutoffs(), categories.get('basepair', []))
    pairs = candidate_pairs(nts)
    prevO3 = previous_O3_map(nts) if 'backbone' in categories else {}
    i2p = defaultdict(list)
    c2i = defaultdict(set)
    u2bp = {}
    tr = {"candidates": [list(p) for p in pairs], "sO": [], "stacking": [], "sugar_ribose": [],
          "backbone": [], "coplanar": [], "basepair": []}
    count_pair = 0
    for i, j in pairs:
        nt1, nt2 = nts[i], nts[j]
        parent1 = get_parent(nt1.sequence)
        parent2 = get_parent(nt2.sequence)
        gly1 = glycosidic(nt1)
        marked = False
        if 'sO' in categories:
            a, ar = check_base_oxygen_stack_rings(nt1, nt2, parent1)
            tr["sO"].append([i, j, a])
            if a:
                count_pair += 1
                i2p[a].append((i, j))
                i2p[ar].append((j, i))
                c2i['sO'].update([a, ar])
            a, ar = check_base_oxygen_stack_rings(nt2, nt1, parent2)
            tr["sO"].append([j, i, a])
            if a:
                count_pair += 1
                i2p[a].append((j, i))
                i2p[ar].append((i, j))
                c2i['sO'].update([a, ar])
        if 'stacking' in categories:
            a, ar = check_base_base_stacking(nt1, nt2, parent1, parent2)
            tr["stacking"].append([i, j, a])
            if a:
                count_pair += 1
                i2p[a].append((i, j))
                c2i['stacking'].update([a, ar])
        if 'sugar_ribose' in categories:
            if parent1 not in ['DA', 'DC', 'DG', 'DT'] and parent2 not in ['DA', 'DC', 'DG', 'DT']:
                a = check_sugar_ribose(nt1, nt2, parent1)
                tr["sugar_ribose"].append([i, j, a])
                if a:
                    count_pair += 1
                    i2p[a].append((i, j))
                    c2i['sugar_ribose'].add(a)
                a = check_sugar_ribose(nt2, nt1, parent2)
                tr["sugar_ribose"].append([j, i, a])
                if a:
                    count_pair += 1
                    i2p[a].append((j, i))
                    c2i['sugar_ribose'].add(a)
        if 'backbone' in categories:
            bph, br = check_base_backbone_interactions(nt1, nt2, prevO3.get(j), parent1)
            tr["backbone"].append([i, j, bph, br])
            if bph:
                count_pair += 1
                i2p[bph].append((i, j))
                c2i['backbone'].add(bph)
            if br:
                count_pair += 1
                i2p[br].append((i, j))
                c2i['backbone'].add(br)
        gly2 = glycosidic(nt2)
        if len(gly2) < 3:
            continue
        res = {}
        for tag, (na, nb, pa, pb, ga, gb) in (("12", (nt1, nt2, parent1, parent2, gly1, gly2)),
                                              ("21", (nt2, nt1, parent2, parent1, gly2, gly1))):
            if (pa or "") + "," + (pb or "") in BASEPAIR_COMBOS:
                pd = {"displ12": np.dot(np.subtract(gb, ga), na.rotation), "parent1": pa, "parent2": pb}
                if 'coplanar' in categories:
                    pd = check_coplanar(na, nb, pd)
                    tr["coplanar"].append([i if tag == "12" else j, j if tag == "12" else i,
                                           bool(pd["coplanar"]), float(pd["gap12"]),
                                           pd.get("min_distance"), pd["coplanar_value"]])
                    if pd["coplanar"] and not marked:
                        count_pair += 1
                        i2p['cp'].append((i, j))
                        c2i['coplanar'].add('cp')
                        marked = True
                LW, sub, q = check_basepair_cutoffs(na, nb, pd, focused[pa + "," + pb],
                                                    HBONDS.get(pa + "," + pb, {}))
                tr["basepair"].append([i if tag == "12" else j, j if tag == "12" else i, LW,
                                       sub if LW else None,
                                       q["cutoff_distance"] if q else None, q["max_gap"] if q else None])
                res[tag] = (LW, sub, q)
            else:
                res[tag] = ("", "", {})
        i12, s12, q12 = res["12"]
        i21, s21, q21 = res["21"]
        i12r = reverse_edges(i12)
        if len(i12) > 0 and len(i21) > 0:                       # NA:935-957
            if i12r.lower() == i21.lower():
                i21 = ""
            elif "n" in i12 and "n" in i21:
                if q12['cutoff_distance'] < q21['cutoff_distance']:
                    i21 = ""
                else:
                    i12 = ""
            elif "n" in i21:
                i21 = ""
            elif "n" in i12:
                i12 = ""
            else:
                i21 = ""
        if len(i12) > 0:
            new = (i12, i12r, q12, i, j)
        elif len(i21) > 0:
            new = (i21, reverse_edges(i21), q21, j, i)
        else:
            new = None
        if new:
            count_pair += 1
            inter, inter_r, q, u1, u2 = new
            u2bp.setdefault(u1, []).append([inter, q, u2])
            u2bp.setdefault(u2, [])
            qr = {'cutoff_distance': q['cutoff_distance'], 'max_gap': q['max_gap'],
                  'atoms1': q['atoms2'], 'atoms2': q['atoms1']}
            u2bp[u2].append([inter_r, qr, u1])
    remove_pairs, make_near_pairs = check_for_two_interactions_on_same_edge(u2bp)
    saved = set()
    for u, bps in u2bp.items():
        for inter, q, v in bps:
            if (u, v) not in remove_pairs and (u, v) not in saved:
                if (u, v) in make_near_pairs:
                    inter = "n" + inter
                i2p[inter].append((u, v))
                saved.add((v, u))
                saved.add((u, v))
                if inter not in c2i['basepair']:
                    r = reverse_edges(inter)
                    c2i['basepair'].update([inter, r])
                    c2i['basepair_detail'].update([inter, r])
    out = calculate_crossing_numbers(nts, i2p)
    annotate_covalent_connections(nts, out, c2i)
    tr["count_pair"] = count_pair
    return out, c2i, tr


# ---------------------------------------------------------------- glycosidic bond orientation (N4)
def bond_orientation(spec, nts=None):
    """annotate_bond_orientation, fr3d/classifiers/NA_unit_annotation.py:48-171: the glycosidic torsion chi
    (O4'-C1'-N9-C4 for purines, O4'-C1'-N1-C2 for pyrimidines, IUPAC sign) and its class per nucleotide.
    Returns the reference's rows: {'unit_id', 'orientation' ('anti' | 'syn' | 'int_syn' | 'NA'), 'chi_degree'
    ('%0.3f' string or None)}."""
    if nts is None:
        nts = build_nts(spec)
    rows = []
    for nt in nts:
        chi = None
        classification = ""
        N1N9 = np.array([])
        C2C4 = np.array([])
        if nt.sequence in ['A', 'G', 'DA', 'DG']:                      # :67-70
            N1N9, C2C4 = nt.center("N9"), nt.center("C4")
        elif nt.sequence in ['C', 'U', 'DC', 'DT']:                    # :71-74
            N1N9, C2C4 = nt.center("N1"), nt.center("C2")
        else:
            parent = get_parent(nt.sequence)                           # :75-95
            p2m = MODIFIED[nt.sequence]["parent_atom_to_modified"] if nt.sequence in MODIFIED else {}
            if parent in ['A', 'G', 'DA', 'DG']:
                if "N9" in p2m:
                    N1N9 = nt.center(p2m["N9"])
                if "C4" in p2m:
                    C2C4 = nt.center(p2m["C4"])
            elif parent in ['C', 'U', 'DC', 'DT']:
                if "N1" in p2m:
                    N1N9 = nt.center(p2m["N1"])
                if "C2" in p2m:
                    C2C4 = nt.center(p2m["C2"])
            else:                                                      # :96-108 no identified parent
                N1N9 = nt.center("N9")
                if len(N1N9) == 3:
                    C2C4 = nt.center("C4")
                else:
                    N1N9, C2C4 = nt.center("N1"), nt.center("C2")
        if len(N1N9) == 3 and len(C2C4) == 3:
            C1P, O4P = nt.center("C1'"), nt.center("O4'")
            if len(C1P) == 3 and len(O4P) == 3:
                O4P_C1P = C1P - O4P                                     # :115-117
                N1N9_C1P = C1P - N1N9
                C2C4_N1N9 = N1N9 - C2C4
                perp_to_sugar = np.cross(O4P_C1P, N1N9_C1P)             # :123-126
                norm_perp_to_sugar = np.linalg.norm(perp_to_sugar)
                if norm_perp_to_sugar != 0:
                    perp_to_sugar = perp_to_sugar / norm_perp_to_sugar
                perp_to_base = np.cross(-N1N9_C1P, C2C4_N1N9)           # :128-130
                perp_to_base = perp_to_base / np.linalg.norm(perp_to_base)
                cross_cross_chi = np.cross(perp_to_base, perp_to_sugar)
                cos_chi = np.dot(perp_to_sugar, perp_to_base)           # :138-142
                if np.dot(cross_cross_chi, N1N9_C1P) > 0:
                    sin_chi = np.linalg.norm(cross_cross_chi)
                else:
                    sin_chi = -np.linalg.norm(cross_cross_chi)
                chi = 180 * math.atan2(sin_chi, cos_chi) / math.pi      # :147
                if chi > -90 and chi < -45:                             # :150-155
                    classification = 'int_syn'
                elif chi >= -45 and chi < 90:
                    classification = 'syn'
                else:
                    classification = 'anti'
        if classification:
            rows.append({'unit_id': nt.unit_id(), 'orientation': classification, 'chi_degree': "%0.3f" % chi,
                         'chi': chi})
        else:
            rows.append({'unit_id': nt.unit_id(), 'orientation': 'NA', 'chi_degree': None, 'chi': None})
    return rows


# ---------------------------------------------------------------- FR3D search screening (N3)
def fixed_radius_search(num_points, models, positions, max_cutoff, square_size=None):
    """fixed_radius_search, fr3d/search/pair_processing.py:106-177: every pair of points of the same model with
    squared distance below max_cutoff^2, as (pnt1, pnt2, distance_squared) in the reference's orientation and list
    order.  The reference finds them through buckets of side max_cutoff keyed by a linearised integer (:133-143;
    square_size only sizes that linearisation - the points span less than it, so distinct cells never collide and
    the keys can be the cell triples themselves); adjacent buckets are the 13 at the same z level or above
    (:147-151), then the bucket itself with pnt1 < pnt2 (:168-174)."""
    d = max_cutoff
    buckets = {}
    for i in range(num_points):
        x, y, z = positions[i]
        if not (np.isnan(x) or np.isnan(y) or np.isnan(z)):
            buckets.setdefault((x // d, y // d, z // d, models[i]), []).append(i)
    dx = [1, 0, -1, 1, 0, -1, 1, 0, -1, 1, 0, -1, 1]
    dy = [1, 1, 1, 0, 0, 0, -1, -1, -1, 1, 1, 1, 0]
    dz = [1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0]
    out = []

    def dist2(a, b):
        pa, pb = positions[a], positions[b]
        return (pa[0] - pb[0]) ** 2 + (pa[1] - pb[1]) ** 2 + (pa[2] - pb[2]) ** 2

    for key, members in buckets.items():
        for j in range(13):
            other = buckets.get((key[0] + dx[j], key[1] + dy[j], key[2] + dz[j], key[3]))
            if other is not None:
                for a in members:
                    for b in other:
                        dd = dist2(a, b)
                        if dd < d * d:
                            out.append((a, b, dd))
        for a in members:
            for b in members:
                if a < b:
                    dd = dist2(a, b)
                    if dd < d * d:
                        out.append((a, b, dd))
    return out


def get_pairlist(Q, models, centers, alt_index=False):
    """get_pairlist, fr3d/search/pair_processing.py:7-104: per pair of query positions, the (index, index) pairs
    (both orders) whose distance lies in that pair's range, sliced out of the distance-sorted triples."""
    pairlist = defaultdict(dict)
    npos = Q['numpositions']
    keys = [(i, j) for i in range(npos) for j in range(i + 1, npos)] if alt_index is False else \
        [(i, j) for a, i in enumerate(alt_index) for j in alt_index[a + 1:]]
    if Q["type"] not in ("mixed", "geometric"):
        for i, j in keys:
            pairlist[i][j] = "full"
        return pairlist
    if centers.size == 0:
        return pairlist
    triples = sorted(fixed_radius_search(len(centers), models, centers, Q["largestMaxRange"]), key=lambda t: t[2])
    both = []
    for a, b, _ in triples:
        both.append((a, b))
        both.append((b, a))
    d2 = [t[2] for t in triples]
    for i, j in keys:
        qi, qj = (i, j) if alt_index is False else (alt_index.index(i), alt_index.index(j))
        lo2 = Q["requireddistanceminimum"][qi][qj] ** 2
        hi2 = Q["requireddistancemaximum"][qi][qj] ** 2
        if alt_index is False:
            # :44-75 - two bisections; the upper one never returns less than index 0 of a non-empty list
            start = 0 if Q["requireddistanceminimum"][qi][qj] <= 0 else sum(1 for v in d2 if v <= lo2)
            end = max(sum(1 for v in d2 if v < hi2) - 1, 0) if d2 else -1
            pairlist[i][j] = both[2 * start:2 * end + 2]
        else:
            rows = []                                                # :78-88
            for a, b, v in triples:
                if v > lo2 and v < hi2:
                    rows.append((alt_index[a], alt_index[b]))
                    rows.append((alt_index[b], alt_index[a]))
            pairlist[i][j] = rows
    return pairlist
