"""Batched base-frame fitting (K1) for host Component objects and raw arrays.

Replaces Component.calculate_rotation_matrix (fr3d/data/components.py:273-349) ->
besttransformation (fr3d/geometry/superpositions.py:10-89) with one CUDA launch for any
number of residues."""

import numpy as np

from .. import packing
from ..engine import DeviceBatch, Engine


def fit_batch_frames(batch, place_h=True, device=None):
    """Run K1 on an NtBatch in place: fills center, rot, hydrogen slots, base_mask; returns status."""
    eng = Engine.get(device)
    torch = eng.torch
    with torch.cuda.device(eng.device):
        db = DeviceBatch(torch, eng.device, batch)
        status = eng.fit_frames(db, place_h=place_h).cpu().numpy()
        batch.center[:] = db.t["center"].cpu().numpy()
        batch.rot[:] = db.t["rot"].cpu().numpy()
        batch.base_xyz[:] = db.t["base_xyz"].cpu().numpy()
        batch.base_mask[:] = db.t["base_mask"].cpu().numpy().view(np.uint32)
    batch.frames_given = True
    return status


def fit_component_frames(components, device=None):
    """Fit every nucleotide Component's frame in one launch and store rotation_matrix
    (numpy.matrix) / base_center on it; residues the kernel rejects keep rotation_matrix None
    (components.py:324-337)."""
    todo = []
    for c in components:
        c.rotation_matrix = None
        c.base_center = None
        if packing._slot_maps(c.sequence) is not None:
            todo.append(c)
    if not todo:
        return
    b = packing.NtBatch(len(todo))
    rows = packing._Rows()
    for c in todo:
        packing._fill_atoms(rows, packing._slot_maps(c.sequence), ((a.name, a.x, a.y, a.z) for a in c._atoms))
    rows.flush(b)
    status = fit_batch_frames(b, place_h=False, device=device)
    for k, c in enumerate(todo):
        if status[k] == 0:
            c._set_frame(b.rot[k].reshape(3, 3), b.center[k])
