"""Write a coordinate spec as mmCIF text (atom_site + the few categories fr3d/cif/reader.py reads: chem_comp, entity,
pdbx_struct_oper_list, pdbx_struct_assembly_gen).  There is no network and the reference ships no CIF file, so the
ingest path is exercised on files written here from the committed specs; coordinates are printed with three
decimals like PDB-distributed files."""

_TYPE = {"DA": "DNA linking", "DC": "DNA linking", "DG": "DNA linking", "DT": "DNA linking"}


def _q(tok):
    """CIF quoting of one value."""
    tok = str(tok)
    if tok == "":
        return "."
    if "'" in tok or '"' in tok or " " in tok or tok.startswith("_") or tok.startswith("#"):
        return '"%s"' % tok if '"' not in tok else "'%s'" % tok
    return tok


def spec_to_cif(spec, operators=None, assembly=None, extra_chem=None, hetero=None):
    """spec: {"pdb", "nts": [...]} (all nts with symmetry '1_555': symmetry copies come from the operators).
    operators: list of {"id", "name", "matrix" 3x3, "vector" 3}; assembly: list of (oper_expression, asym_id_list)
    rows of pdbx_struct_assembly_gen.  label_asym_id = auth_asym_id = chain; label_seq_id = index.  An nt may carry
    "atom_alt": per-atom alternate ids (None = common to all conformers).  hetero: extra residues
    {"chain","sequence","number","index","model","group","entity","atoms"} written after the nucleotides."""
    out = ["data_%s" % spec["pdb"], "#"]
    comps = []
    for nt in spec["nts"]:
        if nt["sequence"] not in comps:
            comps.append(nt["sequence"])
    out += ["loop_", "_chem_comp.id", "_chem_comp.type"]
    for c in comps:
        out.append("%s %s" % (_q(c), _q(_TYPE.get(c, "RNA linking"))))
    for c, t in (extra_chem or {}).items():
        out.append("%s %s" % (_q(c), _q(t)))
    out += ["#", "loop_", "_entity.id", "_entity.type", "1 polymer"]
    ents = {}
    for h in hetero or []:
        ents[h["entity"]] = "polymer" if h["index"] is not None else ("water" if h["sequence"] == "HOH" else "non-polymer")
    for e, t in sorted(ents.items()):
        out.append("%s %s" % (e, t))
    out.append("#")
    if operators:
        cols = ["id", "name"] + ["matrix[%d][%d]" % (r + 1, c + 1) for r in range(3) for c in range(3)] + \
               ["vector[%d]" % (r + 1) for r in range(3)]
        out += ["loop_"] + ["_pdbx_struct_oper_list.%s" % c for c in cols]
        for op in operators:
            vals = [op["id"], op["name"]] + [repr(float(op["matrix"][r][c])) for r in range(3) for c in range(3)] + \
                   [repr(float(op["vector"][r])) for r in range(3)]
            out.append(" ".join(_q(v) for v in vals))
        out.append("#")
    if assembly:
        out += ["loop_", "_pdbx_struct_assembly_gen.assembly_id", "_pdbx_struct_assembly_gen.oper_expression",
                "_pdbx_struct_assembly_gen.asym_id_list"]
        for expr, asyms in assembly:
            out.append("1 %s %s" % (_q(expr), _q(asyms)))
        out.append("#")
    cols = ["group_PDB", "id", "type_symbol", "label_atom_id", "label_alt_id", "label_comp_id", "label_asym_id",
            "label_entity_id", "label_seq_id", "pdbx_PDB_ins_code", "Cartn_x", "Cartn_y", "Cartn_z", "occupancy",
            "B_iso_or_equiv", "pdbx_formal_charge", "auth_seq_id", "auth_comp_id", "auth_asym_id", "auth_atom_id",
            "pdbx_PDB_model_num"]
    out += ["loop_"] + ["_atom_site.%s" % c for c in cols]
    serial = 0
    for nt in list(spec["nts"]) + list(hetero or []):
        alts = nt.get("atom_alt")
        for k, (name, x, y, z) in enumerate(nt["atoms"]):
            serial += 1
            alt = (alts[k] if alts is not None else nt.get("alt_id")) or "."
            vals = [nt.get("group", "ATOM"), serial, name[0], name, alt, nt["sequence"], nt["chain"], nt.get("entity", 1),
                    nt["index"] if nt["index"] is not None else ".", nt.get("insertion_code") or "?",
                    "%.3f" % x, "%.3f" % y, "%.3f" % z, "1.00", "30.00", "?", nt["number"], nt["sequence"], nt["chain"], name,
                    nt["model"]]
            out.append(" ".join(_q(v) for v in vals))
    out.append("#")
    return "\n".join(out) + "\n"
