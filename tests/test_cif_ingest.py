"""mmCIF ingest: fr3d_python_b200.cif.reader against what the UNMODIFIED fr3d.cif.reader produced for
the same CIF text (tests/golden/cif_scenario.cif.gz; tools/make_golden.py cif through tools/pdbx_shim): the same
nucleotides in the same order with the same ids and atoms - DNA and modified residues, alternate conformers with
common atoms, an insertion code, two models, three symmetry operators (one unnamed -> "ASM_3"), hetero / water /
peptide residues filtered out - and therefore byte-equal packed records."""

import gzip
import os

import numpy as np

from conftest import GOLD, load_golden
from fr3d_python_b200 import packing
from fr3d_python_b200.cif import reader, writer


def _text():
    with gzip.open(os.path.join(GOLD, "cif_scenario.cif.gz"), "rt") as f:
        return f.read()


def test_residues_equal_the_reference_readers():
    want = load_golden("cif_scenario_reference_reader.json.gz")
    got = reader.cif_to_spec(_text())
    assert got["pdb"] == want["pdb"] and len(got["nts"]) == len(want["nts"]) == 263
    exact = 0
    for g, w in zip(got["nts"], want["nts"]):
        for k in ("chain", "sequence", "number", "index", "model", "symmetry", "alt_id", "insertion_code"):
            assert g[k] == w[k], (k, g[k], w[k], w["unit_id"])
        assert [a[0] for a in g["atoms"]] == [a[0] for a in w["atoms"]], w["unit_id"]
        ga, wa = np.array([a[1:] for a in g["atoms"]]), np.array([a[1:] for a in w["atoms"]])
        if w["symmetry"] == "ASM_3":          # a general rotation: the reference multiplies atom by atom (BLAS gemv)
            assert np.abs(ga - wa).max() < 1e-12
        else:                                  # identity and axis-flip operators are exact in any summation order
            assert np.array_equal(ga, wa), w["unit_id"]
            exact += 1
    assert exact > 150
    b = packing.pack_specs(got)
    assert b.unit_ids() == [w["unit_id"] for w in want["nts"]]


def test_packed_records_equal_those_of_the_reference_readers_residues():
    want = load_golden("cif_scenario_reference_reader.json.gz")
    a = reader.cif_to_batch(_text())
    b = packing.pack_specs(want)
    for name in ("base_mask", "bb_mask", "meta"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    rot3 = np.array([nt["symmetry"] == "ASM_3" for nt in want["nts"]])
    for name in ("base_xyz", "bb_xyz"):
        x, y = getattr(a, name), getattr(b, name)
        assert np.array_equal(x[~rot3], y[~rot3]), name
        assert np.abs(x[rot3] - y[rot3]).max() < 1e-12
    assert a.chain == b.chain and a.sequence == b.sequence and a.symmetry == b.symmetry and a.alt_id == b.alt_id


def test_writer_reader_round_trip_and_fallbacks(golden):
    """spec -> CIF text -> spec (coordinates rounded to three decimals like a PDB file); a file without assembly
    information gets the reference's fallback operator, which shifts every atom by (1, 1, 1) (reader.py:146-160)."""
    spec = golden("spec_1s72_strand.json.gz")
    ident = [{"id": "1", "name": "1_555", "matrix": [[1, 0, 0], [0, 1, 0], [0, 0, 1]], "vector": [0, 0, 0]}]
    back = reader.cif_to_spec(writer.spec_to_cif(spec, operators=ident, assembly=[("1", "0")]))
    assert len(back["nts"]) == 10
    by_id = {(nt["sequence"], int(nt["number"])): nt for nt in spec["nts"]}
    for nt in back["nts"]:
        src = by_id[(nt["sequence"], nt["number"])]
        assert [a[0] for a in nt["atoms"]] == [a[0] for a in src["atoms"]]
        assert np.abs(np.array([a[1:] for a in nt["atoms"]]) - np.array([a[1:] for a in src["atoms"]])).max() <= 5.0001e-4
    shifted = reader.cif_to_spec(writer.spec_to_cif(spec))
    d = np.array([a[1:] for a in shifted["nts"][0]["atoms"]]) - np.array([a[1:] for a in back["nts"][0]["atoms"]])
    assert np.abs(d - 1.0).max() < 1e-9
    empty = reader.cif_to_spec(writer.spec_to_cif({"pdb": "E", "nts": []}))
    assert empty["nts"] == []


def _same_batch(a, b, tol_rows=None):
    for name in ("base_mask", "bb_mask", "meta"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    for name in ("base_xyz", "bb_xyz"):
        x, y = getattr(a, name), getattr(b, name)
        if tol_rows is None:
            assert np.array_equal(x, y), name
        else:
            assert np.array_equal(x[~tol_rows], y[~tol_rows]), name
            assert np.abs(x[tol_rows] - y[tol_rows]).max() < 1e-12
    assert np.array_equal(a.struct_off, b.struct_off)
    for name in ("pdb", "model", "chain", "sequence", "number", "index", "symmetry", "alt_id", "insertion_code"):
        assert getattr(a, name) == getattr(b, name), name
    assert a.unit_ids() == b.unit_ids()


def test_native_ingest_equals_the_python_reader(golden):
    """csrc/ingest.cpp (C++, one thread per file, no CUDA involved) against cif.reader + packing.pack_specs: every record
    plane, every identity column, the unit ids - on the reference-pinned scenario, on plain files, on a batch of
    several files parsed by several threads, and on a file without assembly information."""
    from fr3d_python_b200.cif import native
    text = _text()
    want = reader.cif_to_batch(text)
    rot3 = np.array([s == "ASM_3" for s in want.symmetry])             # the general rotation: a x + b y + c z + t is
    _same_batch(native.cif_to_batch(text, n_threads=1), want, rot3)    # summed left to right here, by BLAS there
    ident = [{"id": "1", "name": "1_555", "matrix": [[1, 0, 0], [0, 1, 0], [0, 0, 1]], "vector": [0, 0, 0]}]
    texts = [writer.spec_to_cif(golden("spec_1s72_strand.json.gz"), operators=ident, assembly=[("1", "0")]),
             text,
             writer.spec_to_cif(golden("spec_1s72_sarcin.json.gz")),                       # no assembly: (1, 1, 1) shift
             writer.spec_to_cif({"pdb": "EMPTY", "nts": []}),
             writer.spec_to_cif(golden("spec_1gid_full.json.gz"), operators=ident, assembly=[("1", "A,B")])]
    want = reader.cif_to_batch(texts)
    rot3 = np.array([s == "ASM_3" for s in want.symmetry])
    for threads in (1, 3):
        got = native.cif_to_batch(texts, n_threads=threads)
        assert got.n_structures == 5 and got.n == want.n
        _same_batch(got, want, rot3)
    # a quoted value with a blank, a text field and comments in other categories do not derail the atom_site loop
    odd = texts[0].replace("data_1S72\n#", "data_1S72\n#\n_struct.title 'a title with blanks'\n_struct.note\n;free text\n;\n# comment\n")
    _same_batch(native.cif_to_batch(odd, n_threads=1), reader.cif_to_batch(odd))
