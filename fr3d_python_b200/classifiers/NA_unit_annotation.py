"""Glycosidic bond orientation (chi, syn / anti) on the GPU — the reference's entry points of
fr3d/classifiers/NA_unit_annotation.py on top of fr3d_bond_orientation (include/fr3d_b200.h).

    annotate_bond_orientation(structure, pipeline=False)   NA_unit_annotation.py:48-171, same return value
    annotate_bond_orientation_batch(batch)                 every structure of an NtBatch / list of specs, one launch
    write_txt_output_file(...)                             NA_unit_annotation.py:173-191

The torsion and its class are computed on the device for all nucleotides at once; unit ids, the '%0.3f' formatting
of chi and the messages stay on the host.  There is no CPU path.

Deviation: a residue without an identified parent base gets 'NA' (the reference additionally tries the raw atom
names N9 / C4 / N1 / C2 of such a residue, :96-108); the packer keeps no atoms for residues it cannot map.
"""

import ctypes as C
import os

import numpy as np

from .. import _lib
from .. import definitions as defs
from .. import packing
from ..engine import DeviceBatch, Engine

ORIENTATION = ["NA", "anti", "syn", "int_syn"]           # FR3D_ORIENT_* of include/fr3d_b200.h


def _ring_slots():
    """Base slot of C4 (purines) / C2 (pyrimidines) per parent; N9 / N1 is slot 0 by construction."""
    out = (C.c_int32 * _lib.N_PARENTS)()
    for k, parent in enumerate(defs.PARENTS):
        slots = defs.SLOT_OF[parent]
        assert slots["N9" if "N9" in slots else "N1"] == 0
        out[k] = slots["C4"] if "N9" in slots else slots["C2"]
    return out


def bond_orientation_arrays(batch, device=None):
    """(chi_degree float64[n] with NaN where undefined, orientation int32[n]) of every nt of an NtBatch.
    Batches packed from plain coordinates get their frames (and, for modified residues, the ideal-position copies the
    reference averages in) from K1 first."""
    eng = Engine.get(device)
    torch = eng.torch
    with torch.cuda.device(eng.device):
        db = DeviceBatch(torch, eng.device, batch)
        if not batch.frames_given:
            eng.fit_frames(db, place_h=True)
        chi = torch.empty(max(batch.n, 1), dtype=torch.float64, device=eng.device)
        cls = torch.empty(max(batch.n, 1), dtype=torch.int32, device=eng.device)
        _lib.check(eng.lib.fr3d_bond_orientation(C.byref(db.nts), _ring_slots(), chi.data_ptr(), cls.data_ptr(),
                                                 eng.stream_ptr()), "fr3d_bond_orientation")
        eng.launches += 1
        return chi[:batch.n].cpu().numpy(), cls[:batch.n].cpu().numpy()


def _rows(batch, lo, hi, chi, cls, pipeline):
    uid = batch.unit_ids()
    parent = batch.meta[lo:hi, 0].tolist()
    rows, errors = [], []
    for k in range(lo, hi):
        u = uid[k]
        if parent[k - lo] < 0:
            msg = "%s has no identified parent" % u                     # :98-101
            errors.append(msg) if pipeline else print(msg)
        c = int(cls[k])
        if c:
            rows.append({'unit_id': u, 'orientation': ORIENTATION[c], 'chi_degree': "%0.3f" % chi[k]})
        else:
            msg = '%s had a calculation error' % u                      # :162-166
            errors.append(msg) if pipeline else print(msg)
            rows.append({'unit_id': u, 'orientation': 'NA', 'chi_degree': None})
    return rows, errors


def annotate_bond_orientation_batch(batch, pipeline=True, device=None):
    """[(bond_annotations, error_message)] per structure of an NtBatch (or of a list of coordinate specs)."""
    if not isinstance(batch, packing.NtBatch):
        batch = packing.pack_specs(batch)
    chi, cls = bond_orientation_arrays(batch, device)
    return [_rows(batch, int(batch.struct_off[s]), int(batch.struct_off[s + 1]), chi, cls, pipeline)
            for s in range(batch.n_structures)]


def annotate_bond_orientation(structure, pipeline=False):
    """NA_unit_annotation.py:48-171.  Returns (bond_annotations, error_message)."""
    soa = getattr(structure, "_soa_batch", None)
    batch = soa if soa is not None else packing.pack_structures(structure)
    return annotate_bond_orientation_batch(batch, pipeline)[0]


def write_txt_output_file(outputNAPairwiseInteractions, PDBid, bond_annotations, categories):
    """NA_unit_annotation.py:173-191."""
    for category in categories.keys():
        filename = os.path.join(outputNAPairwiseInteractions, PDBid + "_" + category + ".txt")
        with open(filename, 'w') as f:
            if category == 'glycosidic':
                for d in bond_annotations:
                    f.write("%s\t%s\t%s\n" % (d['unit_id'], d['orientation'], d['chi_degree']))
