"""Fixture builders that need the reference tree (/root/reference).

DEV-CONTAINER ONLY.  Nothing under tests/, bench.py or the package imports
this module at run time; it is used by tools/make_golden.py and
tools/export_tables.py to turn the reference's own test data into the plain
"structure spec" JSON fixtures committed under tests/golden/.

A structure spec is a dict:
    {"pdb": str, "nts": [ {"chain","sequence","number","index","model",
                           "symmetry","alt_id","insertion_code",
                           "atoms": [[name,x,y,z], ...]}, ... ]}
"""

import contextlib
import io
import os
import re
import sys

REF = os.environ.get("FR3D_REFERENCE", "/root/reference")


def add_reference_to_path():
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)


# ---------------------------------------------------------------- parsing --

def parse_matlab_fitted(path=None):
    """1GID bases from tests/FR3D-Matlab-1GID-fitted-coordinates.txt
    (SURVEY P1): header '1GID A G103', atom lines 'name x y z'.
    Returns list of (chain, seq, number, [(name,x,y,z)...]) with ALL atoms
    (heavy + Matlab hydrogens)."""
    path = path or os.path.join(REF, "tests", "FR3D-Matlab-1GID-fitted-coordinates.txt")
    out = []
    cur = None
    with open(path) as f:
        for line in f:
            m = re.match(r"^1GID (\w) ([ACGU])(\d+)\s*$", line)
            if m:
                cur = (m.group(1), m.group(2), m.group(3), [])
                out.append(cur)
                continue
            fields = line.split()
            if len(fields) == 4 and cur is not None:
                cur[3].append((fields[0], float(fields[1]), float(fields[2]), float(fields[3])))
    return out


def parse_matlab_rotations(path=None):
    """tests/FR3D-Matlab-1GID-rotation-matrices.txt: blocks of header, centre
    (optionally preceded by '1.0e+002 *'), 3x3 rotation."""
    path = path or os.path.join(REF, "tests", "FR3D-Matlab-1GID-rotation-matrices.txt")
    blocks = []
    with open(path) as f:
        lines = [l.rstrip("\n") for l in f]
    i = 0
    cur = None
    scale = 1.0
    nums = []
    for line in lines:
        m = re.match(r"^1GID (\w) ([ACGU])(\d+)\s*$", line)
        if m:
            if cur is not None:
                blocks.append((cur, nums))
            cur = (m.group(1), m.group(2), m.group(3))
            nums = []
            scale = 1.0
            continue
        ms = re.match(r"^\s*([0-9.]+e[+-]\d+)\s*\*\s*$", line)
        if ms:
            scale = float(ms.group(1))
            continue
        fields = line.split()
        if len(fields) == 3 and cur is not None:
            try:
                vals = [float(x) for x in fields]
            except ValueError:
                continue
            if len(nums) == 0:
                vals = [v * scale for v in vals]
                scale = 1.0
            nums.append(vals)
    if cur is not None:
        blocks.append((cur, nums))
    out = []
    for (chain, seq, num), rows in blocks:
        assert len(rows) == 4, (chain, seq, num, rows)
        out.append({"chain": chain, "sequence": seq, "number": num,
                    "center": rows[0], "rotation": rows[1:4]})
    return out


_ATOM_RE = re.compile(
    r"Atom\(type='(\w+)',\s*name='([^']+)',\s*x=([-0-9.]+),\s*y=([-0-9.]+),\s*z=([-0-9.]+)\)")


def parse_test_components(path):
    """Pull `ntNN_C = Component([Atom(...),...], ... sequence='U', chain='0',
    number='10')` definitions out of a reference test module textually (the
    modules themselves no longer import: SURVEY §0 fact 5)."""
    with open(path) as f:
        text = f.read()
    out = {}
    for m in re.finditer(r"^(nt\d+_\w+) = Component\(\[(.*?)\],(.*?)\)\s*$", text, re.S | re.M):
        name, atoms_txt, kw = m.group(1), m.group(2), m.group(3)
        atoms = [(a[1], float(a[2]), float(a[3]), float(a[4])) for a in _ATOM_RE.findall(atoms_txt)]
        seq = re.search(r"sequence='(\w+)'", kw).group(1)
        chain = re.search(r"chain='(\w+)'", kw).group(1)
        number = re.search(r"number='?(\d+)'?", kw).group(1)
        out[name] = {"sequence": seq, "chain": chain, "number": number, "atoms": atoms}
    return out


def modernise_atom_name(name):
    """Old PDB names -> current: C1* -> C1', O1P -> OP1, O2P -> OP2."""
    name = name.replace("*", "'")
    if name == "O1P":
        return "OP1"
    if name == "O2P":
        return "OP2"
    return name


# ------------------------------------------------------------ spec makers --

BASE_HEAVY = {
    'A': ['N9', 'C4', 'N3', 'N1', 'C6', 'N6', 'C8', 'C5', 'C2', 'N7'],
    'C': ['N1', 'C2', 'O2', 'N3', 'C4', 'N4', 'C6', 'C5'],
    'G': ['N9', 'C4', 'N3', 'N1', 'C6', 'O6', 'C8', 'C5', 'C2', 'N7', 'N2'],
    'U': ['N1', 'C2', 'O2', 'N3', 'C4', 'O4', 'C6', 'C5'],
}


def spec_1gid_base():
    """S_1GID_base: 316 nts, heavy base atoms only, chains A/B,
    index = running count per chain."""
    nts = []
    counter = {}
    for chain, seq, num, atoms in parse_matlab_fitted():
        counter[chain] = counter.get(chain, 0) + 1
        heavy = [[n, x, y, z] for (n, x, y, z) in atoms if n in BASE_HEAVY[seq]]
        nts.append({"chain": chain, "sequence": seq, "number": num, "index": counter[chain],
                    "model": "1", "symmetry": "1_555", "alt_id": None, "insertion_code": None,
                    "atoms": heavy})
    return {"pdb": "1GID", "nts": nts}


def spec_1s72_strand():
    """The ten full-atom 1S72 chain-0 nts (tests/discrepancy_tests.py:23-284),
    pre-inferred H dropped, names modernised, index = number."""
    comps = parse_test_components(os.path.join(REF, "tests", "discrepancy_tests.py"))
    nts = []
    for k in range(10, 20):
        c = comps["nt%d_0" % k]
        atoms = []
        for n, x, y, z in c["atoms"]:
            n = modernise_atom_name(n)
            if n.startswith("H") or n[0].isdigit():
                continue
            atoms.append([n, x, y, z])
        nts.append({"chain": c["chain"], "sequence": c["sequence"], "number": c["number"],
                    "index": int(c["number"]), "model": "1", "symmetry": "1_555",
                    "alt_id": None, "insertion_code": None, "atoms": atoms})
    return {"pdb": "1S72", "nts": nts}


def spec_1s72_sarcin():
    """14 full-atom nts of two sarcin-ricin motifs
    (tests/define_1S72_nucleotides.py:14-51)."""
    comps = parse_test_components(os.path.join(REF, "tests", "define_1S72_nucleotides.py"))
    order = ["nt77_9", "nt78_9", "nt79_9", "nt80_9", "nt102_9", "nt103_9", "nt104_9",
             "nt212_0", "nt213_0", "nt214_0", "nt215_0", "nt225_0", "nt226_0", "nt227_0"]
    nts = []
    for name in order:
        c = comps[name]
        atoms = []
        for n, x, y, z in c["atoms"]:
            n = modernise_atom_name(n)
            if n.startswith("H") or n[0].isdigit():
                continue
            atoms.append([n, x, y, z])
        nts.append({"chain": c["chain"], "sequence": c["sequence"], "number": c["number"],
                    "index": int(c["number"]), "model": "1", "symmetry": "1_555",
                    "alt_id": None, "insertion_code": None, "atoms": atoms})
    return {"pdb": "1S72", "nts": nts}


# ------------------------------------------- reference object construction --

def ref_structure_from_spec(spec):
    """Build fr3d.data.Structure from a spec using the UNMODIFIED reference
    classes (model passed as str, type 'RNA linking': SURVEY §8(c))."""
    add_reference_to_path()
    from fr3d.data import Atom, Component, Structure
    comps = []
    for nt in spec["nts"]:
        seq = nt["sequence"]
        atoms = [Atom(pdb=spec["pdb"], model=nt["model"], chain=nt["chain"], component_id=seq,
                      component_number=nt["number"], component_index=nt["index"],
                      insertion_code=nt["insertion_code"], alt_id=nt["alt_id"],
                      x=a[1], y=a[2], z=a[3], name=a[0], symmetry=nt["symmetry"],
                      polymeric=True, type=a[0][0]) for a in nt["atoms"]]
        ctype = "DNA linking" if seq in ("DA", "DC", "DG", "DT") else "RNA linking"
        with contextlib.redirect_stdout(io.StringIO()):
            comps.append(Component(atoms, pdb=spec["pdb"], model=nt["model"], type=ctype,
                                   chain=nt["chain"], symmetry=nt["symmetry"], sequence=seq,
                                   number=nt["number"], index=nt["index"],
                                   insertion_code=nt["insertion_code"], alt_id=nt["alt_id"],
                                   polymeric=True))
    return Structure(comps, pdb=spec["pdb"])
