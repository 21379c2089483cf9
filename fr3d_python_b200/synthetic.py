"""Synthetic workloads of SURVEY.md §8(d): real RNA coordinates (the 1GID fixture) tiled and
perturbed with fixed PCG64 seeds.  There is no network and the reference ships no CIF file, so
every benchmark / large parity input is generated here from tests/golden/spec_1gid_full.json.gz.

All generators work on packed NtBatch arrays (vectorised numpy, no per-atom Python), and
batch_to_spec() turns any structure of a batch back into the plain coordinate spec the CPU
oracle consumes, so both sides see bit-identical inputs.
"""

import gzip
import json
import os

import numpy as np

from . import _lib
from . import definitions as defs
from . import packing

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_golden_spec(name="spec_1gid_full"):
    with gzip.open(os.path.join(GOLDEN, name + ".json.gz"), "rt") as f:
        return json.load(f)


def random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _observed_only(base):
    """Copy of a packed structure keeping only observed atoms (no frames, no placed hydrogens)."""
    b = packing.NtBatch(base.n)
    b.base_xyz[:] = base.base_xyz
    b.bb_xyz[:] = base.bb_xyz
    b.base_mask[:] = base.base_mask & np.uint32(0x00FFFFFF)
    b.bb_mask[:] = base.bb_mask & np.uint32(~(1 << defs.BB_PREV_O3) & 0xFFFFFFFF)
    b.meta[:] = base.meta
    return b


def _transform(base, rng, sigma, rot=None, shift=None):
    """Rigid motion about the centroid of observed atoms + N(0, sigma) noise on every observed atom
    (applied BEFORE frame fitting)."""
    bm = ((base.base_mask[:, None] >> np.arange(_lib.BASE_SLOTS, dtype=np.uint32)[None, :]) & 1).astype(bool)
    km = ((base.bb_mask[:, None] >> np.arange(_lib.BB_SLOTS, dtype=np.uint32)[None, :]) & 1).astype(bool)
    km[:, defs.BB_PREV_O3] = False
    pts = np.concatenate([base.base_xyz[bm], base.bb_xyz[km]])
    cen = pts.mean(axis=0)
    Q = random_rotation(rng) if rot is None else rot
    t = rng.normal(size=3) * 30.0 if shift is None else shift
    new_base = (base.base_xyz - cen) @ Q.T + cen + t + rng.normal(size=base.base_xyz.shape) * sigma
    new_bb = (base.bb_xyz - cen) @ Q.T + cen + t + rng.normal(size=base.bb_xyz.shape) * sigma
    new_base[~bm] = 0.0
    new_bb[~km] = 0.0
    return new_base, new_bb


def pack_base(spec=None):
    """Packed single structure (observed atoms only) used as the tile."""
    if spec is None:
        spec = load_golden_spec()
    return packing.pack_specs(spec)


def _replicate_identity(out, base, copies, pdb_fn, chain_fn):
    n = base.n
    out.struct_off = None
    out.model = base.model * copies
    out.sequence = base.sequence * copies
    out.number = base.number * copies
    out.index = base.index * copies
    out.symmetry = base.symmetry * copies
    out.alt_id = base.alt_id * copies
    out.insertion_code = base.insertion_code * copies
    chains = []
    for k in range(copies):
        chains.extend(chain_fn(k, c) for c in base.chain)
    out.chain = chains
    return n


def _set_prev_o3(out):
    """previous-O3' slot from the chain topology (same rule as packing._finish_identity, vectorised:
    the tile keeps the base structure's residue order, so 'previous' is the previous row when chain
    and index continue)."""
    n = out.n
    if n < 2:
        return
    grp = out.meta[:, 2]
    ch = out.meta[:, 3]
    idx = out.meta[:, 4]
    cont = np.zeros(n, dtype=bool)
    cont[1:] = (grp[1:] == grp[:-1]) & (ch[1:] == ch[:-1]) & (idx[1:] == idx[:-1] + 1)
    has_o3 = ((out.bb_mask >> 5) & 1).astype(bool)
    ok = cont.copy()
    ok[1:] &= has_o3[:-1]
    rows = np.nonzero(ok)[0]
    out.bb_xyz[rows, defs.BB_PREV_O3] = out.bb_xyz[rows - 1, 5]
    out.bb_mask[rows] |= np.uint32(1 << defs.BB_PREV_O3)


def make_structure_batch(base, n_structures, seed0=4000, sigma=0.10, pdb_prefix="S"):
    """C4: n_structures perturbed copies of `base` (seed = seed0 + k, random rigid motion, sigma A
    noise), one structure each."""
    n = base.n
    out = packing.NtBatch(n * n_structures)
    src = _observed_only(base)
    for k in range(n_structures):
        rng = np.random.Generator(np.random.PCG64(seed0 + k))
        nb, nk = _transform(src, rng, sigma)
        out.base_xyz[k * n:(k + 1) * n] = nb
        out.bb_xyz[k * n:(k + 1) * n] = nk
    out.base_mask[:] = np.tile(src.base_mask, n_structures)
    out.bb_mask[:] = np.tile(src.bb_mask, n_structures)
    out.meta[:] = np.tile(src.meta, (n_structures, 1))
    # one group per (structure, model): the base has group ids 0..g-1
    g = int(src.meta[:, 2].max()) + 1 if n else 1
    out.meta[:, 2] += np.repeat(np.arange(n_structures, dtype=np.int32) * g, n)
    out.meta[:, 5] = np.arange(out.n, dtype=np.int32)
    _replicate_identity(out, base, n_structures, None, lambda k, c: c)
    out.pdb = ["%s%04d" % (pdb_prefix, k) for k in range(n_structures)]
    out.struct_off = np.arange(n_structures + 1, dtype=np.int64) * n
    _set_prev_o3(out)
    return out


def _chain_name(k, c):
    return "%s%d" % (c, k)


def make_tiled_structure(base, copies, lattice, seed, sigma=0.10, pdb="TILE"):
    """C2 / C3: ONE structure made of `copies` rotated, noised copies of `base` on a lattice
    (spacing = bounding box + 6 A), unique chain ids per copy."""
    n = base.n
    src = _observed_only(base)
    bm = ((src.base_mask[:, None] >> np.arange(_lib.BASE_SLOTS, dtype=np.uint32)[None, :]) & 1).astype(bool)
    pts = src.base_xyz[bm]
    span = float((pts.max(axis=0) - pts.min(axis=0)).max()) + 6.0 + 20.0
    rng = np.random.Generator(np.random.PCG64(seed))
    out = packing.NtBatch(n * copies)
    nx, ny, nz = lattice
    assert nx * ny * nz >= copies
    k = 0
    for ix in range(nx):
        for iy in range(ny):
            for iz in range(nz):
                if k >= copies:
                    break
                shift = np.array([ix, iy, iz], dtype=float) * span
                nb, nk = _transform(src, rng, sigma, rot=random_rotation(rng), shift=shift)
                out.base_xyz[k * n:(k + 1) * n] = nb
                out.bb_xyz[k * n:(k + 1) * n] = nk
                k += 1
    out.base_mask[:] = np.tile(src.base_mask, copies)
    out.bb_mask[:] = np.tile(src.bb_mask, copies)
    out.meta[:] = np.tile(src.meta, (copies, 1))
    _replicate_identity(out, base, copies, None, _chain_name)
    # chain ranks must follow Python string order of the new chain ids
    chains = sorted(set(out.chain))
    rank = {c: r for r, c in enumerate(chains)}
    out.meta[:, 3] = np.array([rank[c] for c in out.chain], dtype=np.int32)
    out.meta[:, 5] = np.arange(out.n, dtype=np.int32)
    out.pdb = [pdb]
    out.struct_off = np.array([0, out.n], dtype=np.int64)
    _set_prev_o3(out)
    return out


def batch_to_spec(batch, s=0):
    """Plain coordinate spec of structure s (observed atoms currently flagged in the masks)."""
    lo, hi = int(batch.struct_off[s]), int(batch.struct_off[s + 1])
    nts = []
    for k in range(lo, hi):
        parent = defs.PARENTS[batch.meta[k, 0]]
        atoms = []
        for slot, name in enumerate(defs.BB_NAMES):
            if batch.bb_mask[k] >> slot & 1:
                x = batch.bb_xyz[k, slot]
                atoms.append([name, float(x[0]), float(x[1]), float(x[2])])
        for slot, name in enumerate(defs.SLOT_NAMES[parent]):
            if batch.base_mask[k] >> slot & 1:
                x = batch.base_xyz[k, slot]
                atoms.append([name, float(x[0]), float(x[1]), float(x[2])])
        nts.append({"chain": batch.chain[k], "sequence": batch.sequence[k], "number": batch.number[k],
                    "index": batch.index[k], "model": batch.model[k], "symmetry": batch.symmetry[k],
                    "alt_id": batch.alt_id[k], "insertion_code": batch.insertion_code[k], "atoms": atoms})
    return {"pdb": batch.pdb[s], "nts": nts}


# ------------------------------------------------------------------ C5 ------
def _axis_angle(axis, ang):
    x, y, z = axis
    K = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)


def make_instances(base_c, base_r, N, seed):
    """C5 recipe: instance k = the base motif under a random rigid motion, centres + N(0, 0.5 A),
    each rotation post-multiplied by a random rotation of angle ~N(0, 0.15 rad).
    Identical to tools/make_golden.py:make_instances (tests compare the arrays)."""
    base_c = np.asarray(base_c, dtype=float)
    base_r = np.asarray(base_r, dtype=float)
    rng = np.random.Generator(np.random.PCG64(seed))
    n = base_c.shape[0]
    Cn = np.zeros((N, n, 3))
    Rn = np.zeros((N, n, 3, 3))
    for k in range(N):
        Q = random_rotation(rng)
        t = rng.normal(size=3) * 20.0
        Cn[k] = base_c @ Q.T + t + rng.normal(size=(n, 3)) * 0.5
        for p in range(n):
            axis = rng.normal(size=3)
            axis /= np.linalg.norm(axis)
            ang = rng.normal() * 0.15
            Rn[k, p] = Q @ base_r[p] @ _axis_angle(axis, ang)
    return Cn, Rn
