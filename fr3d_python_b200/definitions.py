"""Rule tables of the annotator, loaded from the exported data files in tables/.

Mirrors the names of fr3d/definitions.py (NAbaseheavyatoms, NAbasehydrogens,
NAbasecoordinates, NAbaseMassiveAndHydrogens) and adds the fixed per-parent
"slot" order the device records use: slot k of parent p is the k-th name of
NAbaseheavyatoms[p] + NAbasehydrogens[p] (definitions.py:230-292).
"""

import ctypes as C
import gzip
import json
import os

from . import _lib

_T = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tables")


def _load(name):
    path = os.path.join(_T, name)
    if name.endswith(".gz"):
        with gzip.open(path, "rt") as f:
            return json.load(f)
    with open(path) as f:
        return json.load(f)


_defs = _load("definitions.json")
NAbaseheavyatoms = _defs["NAbaseheavyatoms"]
NAbasehydrogens = _defs["NAbasehydrogens"]
NAbasecoordinates = _defs["NAbasecoordinates"]
NAbaseMassiveAndHydrogens = _defs["NAbaseMassiveAndHydrogens"]
nt_sugar = _defs["nt_sugar"]
nt_phosphate = _defs["nt_phosphate"]
NAbaseatoms = {s: set(NAbaseheavyatoms[s] + NAbasehydrogens[s]) for s in NAbaseheavyatoms}
PLANES = _load("planes.json")

PARENTS = ['A', 'C', 'G', 'U', 'DT']
PARENT_CODE = {p: k for k, p in enumerate(PARENTS)}
STANDARD_SEQUENCES = ['A', 'C', 'G', 'U', 'DA', 'DC', 'DG', 'DT']
SEQ_TO_PARENT = {'A': 'A', 'C': 'C', 'G': 'G', 'U': 'U', 'DA': 'A', 'DC': 'C', 'DG': 'G', 'DT': 'DT'}

# slot order per parent: heavy atoms then hydrogens; slot 0 is the glycosidic N9 / N1
SLOT_NAMES = {p: list(NAbaseheavyatoms[p]) + list(NAbasehydrogens[p]) for p in PARENTS}
SLOT_OF = {p: {n: k for k, n in enumerate(SLOT_NAMES[p])} for p in PARENTS}
N_HEAVY = {p: len(NAbaseheavyatoms[p]) for p in PARENTS}

# backbone slots (FR3D_BB_SLOTS); slot 9 is the O3' of the previous nucleotide (NA:480-525)
BB_NAMES = ["P", "OP1", "OP2", "O5'", "O4'", "O3'", "O2'", "C1'", "C2'"]
BB_SLOT_OF = {n: k for k, n in enumerate(BB_NAMES)}
BB_PREV_O3 = 9

SO_OXYGENS = ["O2'", "O3'", "O4'", "O5'", "OP1", "OP2"]      # NA:1290

BASEPAIR_COMBOS = ['A,A', 'A,C', 'A,G', 'A,U', 'C,C', 'G,C', 'C,U', 'G,G', 'G,U', 'U,U',
                   'A,DT', 'C,DT', 'G,DT', 'DT,DT']          # NA:654

_modified = None


def modified_table():
    """Modified-nucleotide atom maps (fr3d/data/atom_mappings.txt via fr3d/data/mapping.py)."""
    global _modified
    if _modified is None:
        _modified = _load("modified.json.gz")
    return _modified


def get_parent(sequence):
    """NA:1246-1260."""
    if sequence in SEQ_TO_PARENT:
        return SEQ_TO_PARENT[sequence]
    mod = modified_table()
    if sequence in mod:
        return mod[sequence]["parent"]
    return None


def nt_nt_cutoffs():
    """class_limits_2023.nt_nt_cutoffs as nested dicts in the original insertion order."""
    out = {}
    for combo, ilist in _load("cutoffs.json"):
        out[combo] = {}
        for inter, slist in ilist:
            out[combo][inter] = {}
            for sub, cut in slist:
                out[combo][inter][sub] = cut
    return out


_hbonds = None


def load_ideal_basepair_hydrogen_bonds():
    """Same mapping as fr3d/classifiers/hydrogen_bonds.py:8-112 builds from the CSV:
    {combination: {LW: [(donor, hydrogen, acceptor, '12'|'21'), ...]}}."""
    global _hbonds
    if _hbonds is None:
        raw = _load("hbonds.json")
        _hbonds = {combo: {lw: [tuple(t) for t in lst] for lw, lst in d.items()} for combo, d in raw.items()}
    return _hbonds


# Radii (Angstrom, about the base centre) of the discs the screen kernel uses as necessary conditions: every ring
# polygon and near-ring ellipse of check_base_oxygen_stack_rings lies inside SCREEN_SO_RADIUS[parent], the convex
# hull of check_convex_hull_atoms inside SCREEN_HULL_RADIUS[parent]; tests/test_abi.py proves both for the tables
# that are shipped (the tight values are 2.73 2.07 3.00 2.05 2.01 and 3.50 3.25 3.84 2.79 2.83).
SCREEN_SO_RADIUS = {'A': 2.80, 'C': 2.15, 'G': 3.05, 'U': 2.10, 'DT': 2.10}
SCREEN_HULL_RADIUS = {'A': 3.55, 'C': 3.30, 'G': 3.90, 'U': 2.85, 'DT': 2.90}
# ... and the disc of this radius about the base centre lies INSIDE the hull (inscribed radii 1.956 1.471 1.885
# 1.835 1.492): an atom within it (and |z| < 4.5) overlaps the base without evaluating the half planes.
HULL_INNER_RADIUS = {'A': 1.93, 'C': 1.45, 'G': 1.86, 'U': 1.81, 'DT': 1.47}


def static_tables():
    """Fill the fr3d_static_tables_t host struct."""
    t = _lib.StaticTables()
    for p, parent in enumerate(PARENTS):
        t.so_r2max[p] = SCREEN_SO_RADIUS[parent] ** 2
        t.hull_r2max[p] = SCREEN_HULL_RADIUS[parent] ** 2
        t.hull_r2in[p] = HULL_INNER_RADIUS[parent] ** 2
    for p, parent in enumerate(PARENTS):
        names = SLOT_NAMES[parent]
        # standard coordinates of a DNA/RNA parent: RNA tables for A C G U, DT for DT
        coords = NAbasecoordinates[parent]
        for k, n in enumerate(names):
            for c in range(3):
                t.std_xyz[p][k][c] = coords[n][c]
        t.n_heavy[p] = N_HEAVY[parent]
        t.n_slots[p] = len(names)
        rings = PLANES["rings"][parent]
        t.has_div[p] = 1 if rings["divider"] else 0
        if rings["divider"]:
            for c in range(3):
                t.ring_div[p][c] = rings["divider"][c]
        for k in range(4):
            line = rings["ring5"][k] if k < len(rings["ring5"]) else [0.0, 0.0, 1.0]
            for c in range(3):
                t.ring5[p][k][c] = line[c]
        for k in range(6):
            line = rings["ring6"][k] if k < len(rings["ring6"]) else [0.0, 0.0, 1.0]
            for c in range(3):
                t.ring6[p][k][c] = line[c]
        ell = PLANES["ellipses"][parent]
        for c in range(5):
            t.ell5[p][c] = ell["near5"][c] if ell["near5"] else 0.0
            t.ell6[p][c] = ell["near6"][c]
        hull = PLANES["hull"][parent]
        for k in range(7):
            line = hull[k] if k < len(hull) else [0.0, 0.0, 1.0]
            for c in range(3):
                t.hull[p][k][c] = line[c]
        sites = NAbaseMassiveAndHydrogens[parent]
        t.n_sites[p] = len(sites)
        for s, (h, m, bph, br) in enumerate(sites):
            t.site_h[p][s] = SLOT_OF[parent][h]
            t.site_m[p][s] = SLOT_OF[parent][m]
            t.site_carbon[p][s] = 1 if "C" in m else (0 if "N" in m else -1)
            assert bph[0] == br[0] and bph[1:] == "BPh" and br[1:] == "BR"
            t.site_digit[p][s] = int(bph[0])
        t.sr_slot[p] = SLOT_OF[parent]["N3"] if parent in ("A", "G") else (
            SLOT_OF[parent]["O2"] if parent in ("C", "U") else -1)
    for a, pa in enumerate(PARENTS):
        for b, pb in enumerate(PARENTS):
            t.combo_ok[a * _lib.N_PARENTS + b] = 1 if (pa + "," + pb) in BASEPAIR_COMBOS else 0
    return t
