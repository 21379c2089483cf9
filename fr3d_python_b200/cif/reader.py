"""mmCIF ingest without per-atom Python objects: CIF text -> atom_site columns -> the residues the
reference's reader would build, in its order -> NtBatch.

Restates, on column arrays, what fr3d/cif/reader.py does on the way to `Structure.residues(type=[RNA, DNA linking])`:
  operators and assemblies            reader.py:114-237   (identity operator with translation (1,1,1) when a file has no
                                                           assembly information, :146-160 - SURVEY Q10)
  one Atom per (operator, atom_site row), operator index outermost, coordinates = transform . (x, y, z, 1)
                                      reader.py:486-629
  residues = groups of equal (pdb, model, chain, component_id, component_number, insertion_code, symmetry) after a
  STABLE sort by that key (model, chain, component id compared as strings, number as int)   reader.py:440-451
  alternate conformers: atoms without an alt id are copied into every conformer, conformers sorted by alt id
                                      reader.py:420-438
  residue type from chem_comp         reader.py:259-291, 456-458
Pinned by tests/test_cif_ingest.py against what the UNMODIFIED reader produced for tests/golden/cif_scenario.cif.gz
(tools/make_golden.py cif, through tools/pdbx_shim).  No dependency on `pdbx`.
"""

import re

import numpy as np

from .. import _lib
from .. import packing

_NA_TYPES = ("RNA linking", "DNA linking")


# ------------------------------------------------------------------ tokenizer --
def _tokenize_lines(lines):
    """General mmCIF tokens of a list of lines: (kind, value) with kind in 'data' 'loop' 'name' 'value'."""
    out = []
    k, n_lines = 0, len(lines)
    while k < n_lines:
        line = lines[k]
        if line.startswith(";"):
            parts = [line[1:]]
            k += 1
            while k < n_lines and not lines[k].startswith(";"):
                parts.append(lines[k])
                k += 1
            out.append(("value", "\n".join(parts).strip()))
            k += 1
            continue
        i, n = 0, len(line)
        while i < n:
            c = line[i]
            if c in " \t\r":
                i += 1
            elif c == "#":
                break
            elif c in "'\"":
                j = i + 1
                while j < n and not (line[j] == c and (j + 1 == n or line[j + 1] in " \t\r")):
                    j += 1
                out.append(("value", line[i + 1:j]))
                i = j + 1
            else:
                j = i
                while j < n and line[j] not in " \t\r":
                    j += 1
                tok = line[i:j]
                low = tok.lower()
                if low.startswith("data_"):
                    out.append(("data", tok[5:]))
                elif low == "loop_":
                    out.append(("loop", None))
                elif tok.startswith("_"):
                    out.append(("name", tok))
                else:
                    out.append(("value", tok))
                i = j
        k += 1
    return out


class CifTables(object):
    """The categories of the first data block: name, tables[category] = (columns, rows); atom_site is kept as a 2-D
    numpy array of strings (rows x columns)."""

    def __init__(self, text):
        if isinstance(text, bytes):
            text = text.decode()
        lines = text.split("\n")
        self.name = None
        self.tables = {}
        k, n = 0, len(lines)
        other = []
        while k < n:
            line = lines[k]
            if line.startswith("loop_"):
                j = k + 1
                names = []
                while j < n and lines[j].startswith("_"):
                    names.append(lines[j].strip())
                    j += 1
                if names and names[0].startswith("_atom_site."):
                    e = j                             # the rows: up to the next line that opens something else
                    while e < n and lines[e] and lines[e][0] not in "#_" and not lines[e].startswith("loop_") \
                            and not lines[e].startswith("data_"):
                        e += 1
                    cols = [nm.split(".", 1)[1] for nm in names]
                    self.tables["atom_site"] = (cols, self._atom_rows(lines[j:e], len(cols)))
                    k = e
                    continue
            other.append(line)
            k += 1
        toks = _tokenize_lines(other)
        k = 0
        while k < len(toks):
            kind, val = toks[k]
            if kind == "data":
                if self.name is None:
                    self.name = val
                k += 1
            elif kind == "loop":
                k += 1
                names = []
                while k < len(toks) and toks[k][0] == "name":
                    names.append(toks[k][1])
                    k += 1
                values = []
                while k < len(toks) and toks[k][0] == "value":
                    values.append(toks[k][1])
                    k += 1
                if names:
                    w = len(names)
                    cat = names[0][1:].split(".")[0]
                    self.tables[cat] = ([nm.split(".", 1)[1] for nm in names],
                                        [values[r:r + w] for r in range(0, len(values) - w + 1, w)])
            elif kind == "name":
                cat, attr = val[1:].split(".", 1)
                ent = self.tables.setdefault(cat, ([], [[]]))
                ent[0].append(attr)
                ent[1][0].append(toks[k + 1][1] if k + 1 < len(toks) else "?")
                k += 2
            else:
                k += 1

    @staticmethod
    def _atom_rows(lines, width):
        """atom_site rows as a (rows, width) array of str.  Fast path: whitespace split of the whole block (quoted
        values of atom_site never contain blanks: atom names like "O5'"); general tokenizer otherwise."""
        block = "\n".join(lines)
        toks = block.split()
        if len(toks) % width == 0:
            arr = np.array(toks, dtype=str)
            dq, sq = np.char.startswith(arr, '"'), np.char.startswith(arr, "'")
            quoted = dq | sq
            # a quoted value with a blank inside would show up as a fragment that opens a quote without closing it
            closed = (dq & np.char.endswith(arr, '"')) | (sq & np.char.endswith(arr, "'"))
            if not quoted.any():
                return arr.reshape(-1, width)
            if np.array_equal(quoted, closed & (np.char.str_len(arr) >= 2)):
                arr[quoted] = [t[1:-1] for t in arr[quoted].tolist()]
                return arr.reshape(-1, width)
        vals = [v for kind, v in _tokenize_lines(lines) if kind == "value"]
        return np.array(vals, dtype=str).reshape(-1, width)

    def rows(self, cat):
        cols, rows = self.tables[cat]
        return [dict(zip(cols, r)) for r in rows]


# ------------------------------------------------------------------ operators / assemblies --
def _identity_operator():
    """reader.py:146-160: the fallback operator shifts by (1, 1, 1)."""
    t = np.identity(4)
    t[0:3, 3] = 1.0
    return {"id": "I", "name": "1_555", "transform": t}


def load_operators(tables):
    ops = {}
    if "pdbx_struct_oper_list" in tables.tables:
        for op in tables.rows("pdbx_struct_oper_list"):
            t = np.zeros((4, 4))
            for r in range(3):
                t[r, 3] = float(op["vector[%d]" % (r + 1)])
                for c in range(3):
                    t[r, c] = float(op["matrix[%d][%d]" % (r + 1, c + 1)])
            t[3, 3] = 1.0
            ops[op["id"]] = {"id": op["id"], "name": op.get("name"), "transform": t}
    ident = _identity_operator()
    ops[ident["id"]] = ident
    return ops


def load_assemblies(tables, operators):
    """reader.py:205-237 (the default path; `oldStructures` is empty in the reference tree)."""
    assemblies = {}
    if "pdbx_struct_assembly_gen" not in tables.tables:
        return assemblies
    op = None
    for assembly in tables.rows("pdbx_struct_assembly_gen"):
        opers = assembly["oper_expression"].split(",")
        for asym_id in assembly["asym_id_list"].split(","):
            lst = assemblies.setdefault(asym_id, [])
            for operator in opers:
                if "-" in operator and "X" not in operator and "P" not in operator:
                    a, b = operator.replace("(", "").replace(")", "").split("-")
                    for number in range(int(a), int(b) + 1):
                        op = operators[str(number)]
                        if op["id"] not in [x["id"] for x in lst]:
                            lst.append(op)
                    continue
                elif "X" in operator or "P" in operator:
                    pass                      # the reference applies nothing new here and re-tests the previous `op`
                else:
                    op = operators[operator.replace("(", "").replace(")", "")]
                if op is None:
                    raise _lib.Fr3dError("operator expression %r before any usable operator (the reference fails here too)"
                                         % operator)
                if op["id"] not in [x["id"] for x in lst]:
                    lst.append(op)
    return assemblies


def symmetry_name(op):
    name = op.get("name")
    if not name or name == "?":
        return "ASM_%s" % op["id"]            # reader.py:640-641
    return name


# ------------------------------------------------------------------ residues --
def cif_to_spec(text, pdb=None):
    """The nucleotides of a CIF text as a coordinate spec ({"pdb", "nts": [...]}, the input format of
    packing.pack_specs), exactly the residues - order, ids, atoms, coordinates - that
    fr3d.cif.reader.Cif(handle).structure().residues(type=["RNA linking", "DNA linking"]) yields."""
    tables = text if isinstance(text, CifTables) else CifTables(text)
    pdb = pdb or tables.name
    cols, A = tables.tables["atom_site"]
    col = {c: k for k, c in enumerate(cols)}

    def column(name, default=None):
        if name in col:
            return A[:, col[name]]
        if default is None:
            raise _lib.Fr3dError("atom_site lacks %s" % name)
        return np.full(len(A), default, dtype=str)

    n_rows = len(A)
    chem = {}
    if "chem_comp" in tables.tables:
        for r in tables.rows("chem_comp"):
            chem[r["id"]] = r.get("type")
    else:                                                   # reader.py:266-271 (the inheritance of unknown components
        for b in ("A", "C", "G", "U"):                      # from their chain, :273-289, is not reproduced)
            chem[b] = "RNA linking"
        for b in ("DA", "DC", "DG", "DT"):
            chem[b] = "DNA linking"
    operators = load_operators(tables)
    assemblies = load_assemblies(tables, operators)
    asym = column("label_asym_id")
    comp = column("label_comp_id") if "label_comp_id" in col else column("auth_comp_id")
    name = column("label_atom_id") if "label_atom_id" in col else column("auth_atom_id")
    chain = column("auth_asym_id")
    model = column("pdbx_PDB_model_num", "1")
    ins = column("pdbx_PDB_ins_code", "?")
    alt = column("label_alt_id", ".")
    seq = column("label_seq_id", ".")
    try:
        number = column("auth_seq_id").astype(np.int64)
    except ValueError:                                      # reader.py:607-608
        number = np.array([int(re.sub(r"\D", "", v)) for v in column("auth_seq_id")], dtype=np.int64)
    xyz1 = np.ones((n_rows, 4))
    for k, c in enumerate(("Cartn_x", "Cartn_y", "Cartn_z")):
        xyz1[:, k] = column(c).astype(np.float64)
    # only nucleotide residues matter downstream; filter by residue type first (the reference filters last)
    is_na = np.array([chem.get(c) in _NA_TYPES for c in comp.tolist()], dtype=bool) if n_rows else np.zeros(0, dtype=bool)
    # one atom per (operator index, row), operator index outermost, rows in file order (reader.py:556-585)
    base_rows = np.nonzero(is_na)[0]
    if len(base_rows) == 0:
        return {"pdb": pdb, "nts": []}
    uniq_asym, inv = np.unique(asym[base_rows], return_inverse=True)
    if not assemblies:
        lists = [[_identity_operator()] for _ in uniq_asym]                 # reader.py:546-554
        max_ops = 1
    else:
        lists = [assemblies.get(a) or [_identity_operator()] for a in uniq_asym.tolist()]      # reader.py:653-662
        max_ops = max(len(v) for v in assemblies.values())
    n_ops = np.array([len(l) for l in lists], dtype=np.int64)[inv]
    gen_rows, gen_sym, gen_xyz = [], [], []
    for k in range(max_ops):
        sel = n_ops > k
        rows_k = base_rows[sel]
        inv_k = inv[sel]
        xyz_k = np.empty((len(rows_k), 3))
        sym_k = np.empty(len(rows_k), dtype=object)
        for ai, lst in enumerate(lists):
            if k < len(lst):
                m = inv_k == ai
                if m.any():
                    xyz_k[m] = (xyz1[rows_k[m]] @ lst[k]["transform"].T)[:, :3]       # reader.py:624-629
                    sym_k[m] = symmetry_name(lst[k])
        gen_rows.append(rows_k)
        gen_sym.append(sym_k)
        gen_xyz.append(xyz_k)
    gen_rows = np.concatenate(gen_rows)
    gen_sym = np.concatenate(gen_sym)
    gen_xyz = np.concatenate(gen_xyz)

    # residue key, compared like the reference's tuples: strings as strings, number as int
    def codes(values):
        u, inv = np.unique(values, return_inverse=True)
        return inv
    k_model = codes(model[gen_rows])
    k_chain = codes(chain[gen_rows])
    k_comp = codes(comp[gen_rows])
    k_num = number[gen_rows]
    ins_g = np.where(ins[gen_rows] == "?", "", ins[gen_rows])
    k_ins = codes(ins_g)
    k_sym = codes(gen_sym.astype(str))
    order = np.lexsort((k_sym, k_ins, k_num, k_comp, k_chain, k_model))          # stable, last key is primary
    key = np.stack([k_model, k_chain, k_comp, k_num, k_ins, k_sym], axis=1)[order]
    new_res = np.ones(len(order), dtype=bool)
    new_res[1:] = np.any(key[1:] != key[:-1], axis=1)
    starts = np.nonzero(new_res)[0].tolist() + [len(order)]
    nts = []
    g_rows, g_xyz, g_sym = gen_rows[order], gen_xyz[order], gen_sym[order]
    names_l = name[g_rows].tolist()
    alt_l = alt[g_rows].tolist()
    xyz_l = g_xyz.tolist()
    for a, b in zip(starts[:-1], starts[1:]):
        r0 = int(g_rows[a])
        alts = [None if v == "." else v for v in alt_l[a:b]]
        groups = {}
        for i, v in enumerate(alts):
            groups.setdefault(v, []).append(a + i)
        if len(groups) == 1:
            conformers = [(next(iter(groups)), next(iter(groups.values())))]
        else:                                               # reader.py:420-438
            common = groups.pop(None, [])
            conformers = [(v, idx + common) for v, idx in groups.items()]
            conformers.sort(key=lambda t: t[0])
        seq_v = seq[r0]
        for alt_id, idx in conformers:
            nts.append({"chain": str(chain[r0]), "sequence": str(comp[r0]), "number": int(number[r0]),
                        "index": int(seq_v) if seq_v not in (".", "", "?") else None, "model": str(model[r0]),
                        "symmetry": str(g_sym[a]), "alt_id": alt_id, "insertion_code": None if ins[r0] == "?" else str(ins[r0]),
                        "atoms": [[names_l[i]] + xyz_l[i] for i in idx]})
    return {"pdb": pdb, "nts": nts}


def cif_to_batch(texts):
    """One NtBatch for a list of CIF texts (or one text): cif_to_spec + packing.pack_specs."""
    if isinstance(texts, (str, bytes, CifTables)):
        texts = [texts]
    return packing.pack_specs([cif_to_spec(t) for t in texts])
