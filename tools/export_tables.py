"""Export the reference's DATA tables into fr3d_python_b200/tables/*.json.

DEV-CONTAINER ONLY (needs /root/reference).  The product reads the JSON
files; it never imports the reference.  What is exported is data, not code:

  definitions.json   standard base coordinates, heavy/hydrogen atom lists,
                     base-backbone donor sites      (fr3d/definitions.py:230-611)
  cutoffs.json       basepair cutoff boxes, dict order preserved
                     (fr3d/classifiers/class_limits_2023.py)
  hbonds.json        ideal base-pair hydrogen bond atoms
                     (fr3d/classifiers/hydrogen_bonds.py:8-112 applied to the CSV)
  planes.json        ring / convex-hull half planes and near-ring ellipses,
                     the numeric literals of
                     NA_pairwise_interactions.py:1316-1368, 1412-1435, 1483-1524
  modified.json.gz   modified-nucleotide atom maps (fr3d/data/atom_mappings.txt)

Run:  python tools/export_tables.py
"""

import gzip
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_fixtures as rf  # noqa: E402

rf.add_reference_to_path()
OUT = os.path.join(os.path.dirname(HERE), "fr3d_python_b200", "tables")

NUM = r"(-?\s*\d+\.\d+)"


def _f(s):
    return float(s.replace(" ", ""))


def export_definitions():
    from fr3d import definitions as defs
    seqs = ['A', 'C', 'G', 'U', 'DA', 'DC', 'DG', 'DT']
    d = {
        "NAbaseheavyatoms": {s: list(defs.NAbaseheavyatoms[s]) for s in seqs},
        "NAbasehydrogens": {s: list(defs.NAbasehydrogens[s]) for s in seqs},
        "NAbasecoordinates": {s: {k: [float(x) for x in v] for k, v in defs.NAbasecoordinates[s].items()}
                              for s in seqs},
        "NAbaseMassiveAndHydrogens": {s: [list(x) for x in defs.NAbaseMassiveAndHydrogens[s]]
                                      for s in ['A', 'C', 'G', 'U', 'DT', 'DA', 'DC', 'DG']},
        "nt_sugar": {s: list(defs.nt_sugar[s]) for s in seqs},
        "nt_phosphate": {s: list(defs.nt_phosphate[s]) for s in seqs},
    }
    return d


def export_cutoffs():
    from fr3d.classifiers.class_limits_2023 import nt_nt_cutoffs
    out = []
    # list-of-lists keeps dict insertion order explicit (it decides tie-breaks,
    # NA_pairwise_interactions.py:2546-2547, 2682-2686)
    for combo, inters in nt_nt_cutoffs.items():
        ilist = []
        for inter, subs in inters.items():
            slist = []
            for sub, cut in subs.items():
                slist.append([sub, {k: float(v) for k, v in cut.items()}])
            ilist.append([inter, slist])
        out.append([combo, ilist])
    return out


def export_hbonds():
    from fr3d.classifiers.hydrogen_bonds import load_ideal_basepair_hydrogen_bonds
    hb = load_ideal_basepair_hydrogen_bonds()
    return {combo: {lw: [list(t) for t in lst] for lw, lst in d.items()} for combo, d in hb.items()}


def _function_text(src, name):
    m = re.search(r"^def %s\(.*?(?=^def )" % name, src, re.S | re.M)
    return m.group(0)


LINE_RE = re.compile(r"if\s+" + NUM + r"\*x \+\s+" + NUM + r"\*y \+\s+" + NUM + r" > 0:")
ELL_RE = re.compile(r"if " + NUM + r"\*\(x-\(" + NUM + r"\)\)\*\*2 \+ " + NUM + r"\*\(x-\(" + NUM +
                    r"\)\)\*\(y-\(" + NUM + r"\)\) \+ \(y-\(" + NUM + r"\)\)\*\*2 < " + NUM)


def _split_by_parent(text, var):
    """Split a rule body into {parent: text} using the `parentX == 'A'` tests."""
    pieces = {}
    marks = [(m.start(), m.group(1)) for m in re.finditer(r"(?:if|elif) %s == '(\w+)'" % var, text)]
    for k, (pos, parent) in enumerate(marks):
        end = marks[k + 1][0] if k + 1 < len(marks) else len(text)
        pieces[parent] = text[pos:end]
    return pieces


def export_planes():
    path = os.path.join(rf.REF, "fr3d", "classifiers", "NA_pairwise_interactions.py")
    with open(path) as f:
        src = f.read()
    so = _function_text(src, "check_base_oxygen_stack_rings")
    ring_part, ell_part = so.split("# check ellipses only")
    rings = {}
    for parent, txt in _split_by_parent(ring_part, "parent1").items():
        lines = [[_f(a), _f(b), _f(c)] for a, b, c in LINE_RE.findall(txt)]
        if parent in ("A", "G"):
            assert len(lines) == 10, (parent, len(lines))
            rings[parent] = {"divider": lines[0], "ring5": lines[1:5], "ring6": lines[5:10]}
        else:
            assert len(lines) == 6, (parent, len(lines))
            rings[parent] = {"divider": None, "ring5": [], "ring6": lines}
    ellipses = {}
    for parent, txt in _split_by_parent(ell_part, "parent1").items():
        ells = [[_f(v) for v in e] for e in ELL_RE.findall(txt)]
        # each: a, cx, b, cx, cy, cy, r  ->  keep a, b, cx, cy, r
        ells = [[e[0], e[2], e[1], e[4], e[6]] for e in ells]
        if parent in ("A", "G"):
            assert len(ells) == 2
            ellipses[parent] = {"near5": ells[0], "near6": ells[1]}
        else:
            assert len(ells) == 1
            ellipses[parent] = {"near5": None, "near6": ells[0]}
    hull_txt = _function_text(src, "check_convex_hull_atoms")
    hull = {}
    for parent, txt in _split_by_parent(hull_txt, "parent").items():
        lines = [[_f(a), _f(b), _f(c)] for a, b, c in LINE_RE.findall(txt)]
        assert 5 <= len(lines) <= 7, (parent, len(lines))
        hull[parent] = lines
    assert set(rings) == set(ellipses) == set(hull) == {"A", "C", "G", "U", "DT"}
    return {"rings": rings, "ellipses": ellipses, "hull": hull}


def export_modified():
    from fr3d.data import mapping as mp
    out = {}
    for mod, parent in mp.modified_base_to_parent.items():
        out[mod] = {
            "parent": parent,
            "parent_atom_to_modified": dict(mp.parent_atom_to_modified[mod]),
            "modified_atom_to_parent": dict(mp.modified_atom_to_parent[mod]),
            "base_atoms": list(mp.modified_base_atom_list[mod]),
        }
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, fn in [("definitions", export_definitions), ("cutoffs", export_cutoffs),
                     ("hbonds", export_hbonds), ("planes", export_planes)]:
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump(fn(), f, indent=0, separators=(",", ":"))
        print("wrote", name)
    with gzip.GzipFile(os.path.join(OUT, "modified.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(export_modified(), separators=(",", ":")).encode())
    print("wrote modified")


if __name__ == "__main__":
    main()
