"""Host packing: structures -> structure-of-arrays nucleotide records (NtBatch).

This flattens what the annotator reads from fr3d.data.Component /
AtomProxy (fr3d/data/components.py:118-202, fr3d/data/base.py:80-209) into the
per-nt device record documented in include/fr3d_b200.h.  Two sources:

  pack_specs(specs)            plain coordinate specs (what a CIF atom_site loop gives);
                               frames and hydrogens are then produced on the GPU (K1)
  pack_structures(structures)  duck-typed fr3d Structure/Component objects that already
                               carry rotation_matrix / centers / inferred hydrogens
"""

import numpy as np

from . import _lib
from . import definitions as defs

_AMINO = {'A': ("N7", "H61", "H62"), 'C': ("C5", "H41", "H42"), 'G': ("N1", "H21", "H22")}  # components.py:410-428


class NtBatch(object):
    """SoA nucleotide records for one or more structures (host side, numpy)."""

    def __init__(self, n):
        self.n = n
        self.center = np.zeros((n, 3))
        self.rot = np.zeros((n, 9))
        self.base_xyz = np.zeros((n, _lib.BASE_SLOTS, 3))
        self.bb_xyz = np.zeros((n, _lib.BB_SLOTS, 3))
        self.base_mask = np.zeros(n, dtype=np.uint32)
        self.bb_mask = np.zeros(n, dtype=np.uint32)
        self.meta = np.zeros((n, _lib.META_INTS), dtype=np.int32)
        self.frames_given = False
        # identity (host only)
        self.struct_off = np.zeros(1, dtype=np.int64)
        self.pdb = []            # per structure
        self.model = []          # per nt (str)
        self.chain = []
        self.sequence = []
        self.number = []
        self.index = []
        self.symmetry = []
        self.alt_id = []
        self.insertion_code = []
        self._unit_ids = None
        self._unit_id_fn = None

    @property
    def n_structures(self):
        return len(self.struct_off) - 1

    def structure_of(self):
        """int array [n]: structure index of every nt."""
        return np.repeat(np.arange(self.n_structures), np.diff(self.struct_off))

    def unit_ids(self):
        if self._unit_ids is None:
            if self._unit_id_fn is not None:
                self._unit_ids = self._unit_id_fn()
            else:
                sof = self.structure_of()
                self._unit_ids = [encode_unit_id(self.pdb[sof[k]], self.model[k], self.chain[k], self.sequence[k],
                                                 self.number[k], self.alt_id[k], self.insertion_code[k],
                                                 self.symmetry[k]) for k in range(self.n)]
        return self._unit_ids

    def nbytes_device(self):
        return sum(a.nbytes for a in (self.center, self.rot, self.base_xyz, self.bb_xyz, self.base_mask,
                                      self.bb_mask, self.meta))


def select_structures(batch, which):
    """Sub-batch holding the structures listed in `which` (in that order); nt-level arrays are
    sliced, group ids are kept (they only need to be distinct per structure/model)."""
    which = [int(w) for w in which]
    rows = np.concatenate([np.arange(batch.struct_off[w], batch.struct_off[w + 1]) for w in which]) \
        if which else np.zeros(0, dtype=np.int64)
    out = NtBatch(len(rows))
    for name in ("center", "rot", "base_xyz", "bb_xyz", "base_mask", "bb_mask", "meta"):
        getattr(out, name)[:] = getattr(batch, name)[rows]
    out.frames_given = batch.frames_given
    out.pdb = [batch.pdb[w] for w in which]
    lens = [int(batch.struct_off[w + 1] - batch.struct_off[w]) for w in which]
    out.struct_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    for name in ("model", "chain", "sequence", "number", "index", "symmetry", "alt_id", "insertion_code"):
        src = getattr(batch, name)
        dst = []
        for w in which:
            dst.extend(src[int(batch.struct_off[w]):int(batch.struct_off[w + 1])])
        setattr(out, name, dst)
    if batch._unit_ids is not None:
        out._unit_ids = [batch._unit_ids[r] for r in rows]
    return out


def encode_unit_id(pdb, model, chain, sequence, number, alt_id, insertion_code, symmetry):
    """fr3d/unit_ids.py:31-64 for a component (no atom name)."""
    ordered = ['' if v is None else str(v) for v in
               (pdb, model, chain, sequence, number, None, alt_id, insertion_code, symmetry)]
    if ordered[8] == '1_555':
        ordered.pop()
    while ordered and not ordered[-1]:
        ordered.pop()
    return '|'.join(ordered)


_slot_cache = {}


def _slot_maps(sequence):
    """(parent_code, flags, base name->slot, backbone name->slot, dup_mask) for a residue name, or None.
    dup_mask: slots of a modified residue whose atom is in the reference's mapping, i.e. the slots that
    receive an ideal-position copy when no hydrogen is observed (components.py:505-513)."""
    if sequence in _slot_cache:
        return _slot_cache[sequence]
    parent = defs.get_parent(sequence)
    if parent is None:
        res = None
    elif sequence in defs.SEQ_TO_PARENT:
        flags = _lib.FLAG_STD_RNA if sequence in ('A', 'C', 'G', 'U') else 0
        res = (defs.PARENT_CODE[parent], flags, defs.SLOT_OF[parent], defs.BB_SLOT_OF, 0)
    else:
        # modified residue: parent atom X is read through parent_atom_to_modified (NA:324-371)
        p2m = defs.modified_table()[sequence]["parent_atom_to_modified"]
        base = {p2m.get(x, x): k for x, k in defs.SLOT_OF[parent].items()}
        bb = {p2m.get(x, x): k for x, k in defs.BB_SLOT_OF.items()}
        dup = 0
        for x, k in defs.SLOT_OF[parent].items():
            if x in p2m:
                dup |= 1 << k
        res = (defs.PARENT_CODE[parent], _lib.FLAG_MODIFIED, base, bb, dup)
    _slot_cache[sequence] = res
    return res


def _fill_atoms(batch, k, maps, atoms, frames_given=False):
    """atoms: iterable of (name, x, y, z).  Duplicate names are averaged like AtomProxy
    (fr3d/data/base.py:150-166).  Modified residues: see FR3D_FLAG_DUP in include/fr3d_b200.h."""
    pcode, flags, bslot, bbslot, dupmask = maps
    bcount = {}
    saw_h = False
    for name, x, y, z in atoms:
        if 'H' in name:
            saw_h = True
        s = bslot.get(name)
        if s is not None:
            if batch.base_mask[k] >> s & 1:
                c = bcount.get(s, 1)
                batch.base_xyz[k, s] = (batch.base_xyz[k, s] * c + (x, y, z)) / (c + 1)
                bcount[s] = c + 1
            else:
                batch.base_xyz[k, s] = (x, y, z)
                batch.base_mask[k] |= np.uint32(1 << s)
            continue
        s = bbslot.get(name)
        if s is not None:
            if batch.bb_mask[k] >> s & 1:
                c = bcount.get(100 + s, 1)
                batch.bb_xyz[k, s] = (batch.bb_xyz[k, s] * c + (x, y, z)) / (c + 1)
                bcount[100 + s] = c + 1
            else:
                batch.bb_xyz[k, s] = (x, y, z)
                batch.bb_mask[k] |= np.uint32(1 << s)
    if flags & _lib.FLAG_MODIFIED:
        if frames_given:
            # Component objects of the reference already hold the ideal copies: twin slots are those seen twice
            twin = 0
            for s_, c in bcount.items():
                if s_ < 100 and c == 2:
                    twin |= 1 << s_
            if twin:
                flags |= _lib.FLAG_DUP
                batch.meta[k, 7] = twin << 16
        elif not saw_h:
            flags |= _lib.FLAG_DUP
            batch.meta[k, 7] = dupmask
    batch.meta[k, 0] = pcode
    batch.meta[k, 1] = flags
    batch.base_mask[k] |= np.uint32(((pcode + 1) & 15) << 16 | (flags & 15) << 20)   # read by the classify kernel


def _fix_amino_hydrogens(batch, k, parent):
    """components.py:405-445: observed amino hydrogens are relabelled so that H62 / H42 / H22 is
    the one nearer N7 / C5 / N1."""
    if parent not in _AMINO:
        return
    heavy, h1, h2 = _AMINO[parent]
    so = defs.SLOT_OF[parent]
    sh, s1, s2 = so[heavy], so[h1], so[h2]
    m = int(batch.base_mask[k])
    if (m >> sh & 1) and (m >> s1 & 1) and (m >> s2 & 1):
        d1 = np.sum((batch.base_xyz[k, sh] - batch.base_xyz[k, s1]) ** 2)
        d2 = np.sum((batch.base_xyz[k, sh] - batch.base_xyz[k, s2]) ** 2)
        if d1 < d2:
            tmp = batch.base_xyz[k, s1].copy()
            batch.base_xyz[k, s1] = batch.base_xyz[k, s2]
            batch.base_xyz[k, s2] = tmp


def _finish_identity(batch):
    """group ids, chain ranks, alt-id keys and previous-O3' (map_unit_id_to_previous_O3, NA:480-525)."""
    n = batch.n
    sof = batch.structure_of() if n else np.zeros(0, dtype=np.int64)
    groups = {}
    chains = sorted(set(batch.chain))
    chain_rank = {c: r for r, c in enumerate(chains)}
    any_alt = any(a is not None for a in batch.alt_id)
    resid = {}
    alts = {}
    for k in range(n):
        g = groups.setdefault((int(sof[k]), batch.model[k]), len(groups))
        batch.meta[k, 2] = g
        batch.meta[k, 3] = chain_rank[batch.chain[k]]
        idx = batch.index[k]
        batch.meta[k, 4] = 0 if idx is None else int(idx)
        if any_alt:     # alt-id twins screen, NA:706-713
            key = (int(sof[k]), batch.number[k], batch.sequence[k], batch.chain[k], batch.symmetry[k],
                   batch.insertion_code[k])
            batch.meta[k, 5] = resid.setdefault(key, len(resid))
            batch.meta[k, 6] = alts.setdefault(batch.alt_id[k], len(alts))
        else:
            batch.meta[k, 5] = k
            batch.meta[k, 6] = 0
    # previous O3'
    uid = batch.unit_ids()
    for s in range(batch.n_structures):
        lo, hi = int(batch.struct_off[s]), int(batch.struct_off[s + 1])
        order = sorted(range(lo, hi), key=lambda k: (batch.model[k], batch.symmetry[k] or "", batch.chain[k],
                                                     batch.meta[k, 4], uid[k]))
        prev = None
        for k in order:
            if prev is not None and batch.model[k] == batch.model[prev] and batch.symmetry[k] == batch.symmetry[prev] \
                    and batch.chain[k] == batch.chain[prev] and batch.meta[k, 4] == batch.meta[prev, 4] + 1:
                if batch.bb_mask[prev] >> 5 & 1:     # O3' of the previous nt
                    batch.bb_xyz[k, defs.BB_PREV_O3] = batch.bb_xyz[prev, 5]
                    batch.bb_mask[k] |= np.uint32(1 << defs.BB_PREV_O3)
            prev = k


def pack_specs(specs):
    """specs: list of {"pdb", "nts": [{"chain","sequence","number","index","model","symmetry",
    "alt_id","insertion_code","atoms":[[name,x,y,z],...]}]}.  Only RNA/DNA residues belong in a spec."""
    if isinstance(specs, dict):
        specs = [specs]
    n = sum(len(s["nts"]) for s in specs)
    b = NtBatch(n)
    off = [0]
    k = 0
    for spec in specs:
        b.pdb.append(spec["pdb"])
        for nt in spec["nts"]:
            seq = nt["sequence"]
            b.model.append(nt["model"]); b.chain.append(nt["chain"]); b.sequence.append(seq)
            b.number.append(nt["number"]); b.index.append(nt["index"]); b.symmetry.append(nt["symmetry"])
            b.alt_id.append(nt.get("alt_id")); b.insertion_code.append(nt.get("insertion_code"))
            maps = _slot_maps(seq)
            if maps is None:
                b.meta[k, 0] = -1
            else:
                _fill_atoms(b, k, maps, nt["atoms"])
                if not (maps[1] & _lib.FLAG_MODIFIED):
                    _fix_amino_hydrogens(b, k, defs.PARENTS[maps[0]])
            k += 1
        off.append(k)
    b.struct_off = np.array(off, dtype=np.int64)
    _finish_identity(b)
    return b


def pack_structures(structures, chains=None):
    """Duck-typed fr3d Structures (reference objects or fr3d_python_b200.data ones).  Reads, per
    residue of type RNA/DNA linking: rotation_matrix, centers['base'], atoms(), ids — the same
    attributes annotate_nt_nt_in_structure touches (NA:1226-1229, 661-724)."""
    if not isinstance(structures, (list, tuple)):
        structures = [structures]
    comps_per = []
    for st in structures:
        if chains:
            bases = st.residues(chain=chains, type=["RNA linking", "DNA linking"])
        else:
            bases = st.residues(type=["RNA linking", "DNA linking"])
        comps_per.append(list(bases))
    n = sum(len(c) for c in comps_per)
    b = NtBatch(n)
    b.frames_given = True
    off = [0]
    k = 0
    uids = []
    for st, comps in zip(structures, comps_per):
        b.pdb.append(getattr(st, "pdb", None))
        for c in comps:
            seq = c.sequence
            b.model.append(c.model); b.chain.append(c.chain); b.sequence.append(seq)
            b.number.append(c.number); b.index.append(c.index); b.symmetry.append(c.symmetry)
            b.alt_id.append(c.alt_id); b.insertion_code.append(c.insertion_code)
            uids.append(c.unit_id())
            maps = _slot_maps(seq)
            if maps is None:
                b.meta[k, 0] = -1
            else:
                _fill_atoms(b, k, maps, ((a.name, a.x, a.y, a.z) for a in c.atoms()), frames_given=True)
                R = c.rotation_matrix
                ctr = c.centers["base"]
                if R is not None and len(ctr) == 3:
                    b.rot[k] = np.asarray(R, dtype=float).reshape(9)
                    b.center[k] = np.asarray(ctr, dtype=float)
                    b.base_mask[k] |= np.uint32(_lib.MASK_HAS_FRAME)
            k += 1
        off.append(k)
    b.struct_off = np.array(off, dtype=np.int64)
    b._unit_ids = uids
    _finish_identity(b)
    return b
