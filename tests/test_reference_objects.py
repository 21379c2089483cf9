"""Duck-typing boundary: pack_structures() fed with the REFERENCE's own fr3d.data objects.
Runs only where /root/reference is mounted (the dev container); skipped on the GPU box."""

import contextlib
import io
import os
import sys

import numpy as np
import pytest

REF = os.environ.get("FR3D_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "fr3d")), reason="reference tree not mounted")


def _ref_structure(spec):
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from fr3d.data import Atom, Component, Structure
    comps = []
    for nt in spec["nts"]:
        atoms = [Atom(pdb=spec["pdb"], model=nt["model"], chain=nt["chain"], component_id=nt["sequence"],
                      component_number=nt["number"], component_index=nt["index"], x=a[1], y=a[2], z=a[3], name=a[0],
                      symmetry=nt["symmetry"], polymeric=True, type=a[0][0]) for a in nt["atoms"]]
        with contextlib.redirect_stdout(io.StringIO()):
            comps.append(Component(atoms, pdb=spec["pdb"], model=nt["model"], type="RNA linking", chain=nt["chain"],
                                   symmetry=nt["symmetry"], sequence=nt["sequence"], number=nt["number"],
                                   index=nt["index"], polymeric=True))
    return Structure(comps, pdb=spec["pdb"])


def test_reference_structure_packs_to_the_same_records(golden):
    from fr3d_python_b200 import _lib, packing
    from fr3d_python_b200 import definitions as defs
    spec = golden("spec_1s72_strand.json.gz")
    frames = golden("frames_1s72_strand.json.gz")
    st = _ref_structure(spec)
    b = packing.pack_structures(st)
    raw = packing.pack_specs(spec)
    assert b.frames_given and b.n == raw.n == 10
    assert b.unit_ids() == [f["unit_id"] for f in frames]
    for k, f in enumerate(frames):
        assert np.array_equal(b.center[k], f["center"])                     # reference's own frame, untouched
        assert np.array_equal(b.rot[k].reshape(3, 3), f["rotation"])
        assert b.base_mask[k] & _lib.MASK_HAS_FRAME
        parent = defs.PARENTS[b.meta[k, 0]]
        nslots = len(defs.SLOT_NAMES[parent])
        assert b.base_mask[k] & 0xFFFF == (1 << nslots) - 1                 # heavy atoms + the reference's hydrogens
        for name, xyz in f["hydrogens"].items():
            assert np.array_equal(b.base_xyz[k, defs.SLOT_OF[parent][name]], xyz)
    heavy = raw.base_mask & 0xFFFF
    for k in range(b.n):
        for s in range(16):
            if heavy[k] >> s & 1:
                assert np.array_equal(b.base_xyz[k, s], raw.base_xyz[k, s])
    assert np.array_equal(b.bb_xyz, raw.bb_xyz) and np.array_equal(b.bb_mask, raw.bb_mask)
    assert np.array_equal(b.meta[:, [0, 1, 3, 4]], raw.meta[:, [0, 1, 3, 4]])
    # chain filter goes through the reference's EntitySelector
    assert packing.pack_structures(st, chains=["0"]).n == 10 and packing.pack_structures(st, chains=["X"]).n == 0


def test_box_tables_equal_the_reference_focus():
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        from fr3d.classifiers import NA_pairwise_interactions as RNA
    from fr3d_python_b200.classifiers import NA_pairwise_interactions as NA
    lw = NA.Leontis_Westhof_basepairs
    assert NA.focus_basepair_cutoffs(NA.nt_nt_cutoffs, lw) == RNA.focus_basepair_cutoffs(RNA.nt_nt_cutoffs, lw)
    assert NA.nt_nt_cutoffs == RNA.nt_nt_cutoffs
    ref_hb = RNA.load_ideal_basepair_hydrogen_bonds()
    assert NA.load_ideal_basepair_hydrogen_bonds() == ref_hb
    for inter in ["cWW", "tHS", "ntHS", "cWSa", "ncWSa", "s35", "cp", "cSR"]:
        assert NA.reverse_edges(inter) == RNA.reverse_edges(inter)
    for seq in ["A", "DG", "DT", "PSU", "5MC", "OMG", "1MA", "XYZ"]:
        assert NA.get_parent(seq) == RNA.get_parent(seq)


def test_reference_modified_components_pack_with_twin_masks(golden):
    """The reference's Components of modified residues hold ideal-position duplicates of their base atoms;
    pack_structures must flag exactly the slots that are both observed and copied."""
    from fr3d_python_b200 import _lib, packing
    spec = golden("spec_1gid_modified.json.gz")
    sub = {"pdb": spec["pdb"], "nts": spec["nts"][:60]}
    st = _ref_structure(sub)
    b = packing.pack_structures(st)
    raw = packing.pack_specs(sub)
    n_mod = 0
    for k in range(b.n):
        if raw.meta[k, 1] & _lib.FLAG_MODIFIED:
            n_mod += 1
            dup = int(raw.meta[k, 7]) & 0xFFFF if raw.meta[k, 1] & _lib.FLAG_DUP else 0
            observed = int(raw.base_mask[k]) & 0xFFFF
            assert (int(b.meta[k, 7]) >> 16) & 0xFFFF == dup & observed
            assert int(b.base_mask[k]) & 0xFFFF == observed | dup
            assert bool(b.meta[k, 1] & _lib.FLAG_DUP) == bool(dup & observed)
        else:
            assert b.meta[k, 7] == 0
    assert n_mod >= 10
