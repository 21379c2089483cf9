"""Host packing: structures -> structure-of-arrays nucleotide records (NtBatch).

This flattens what the annotator reads from fr3d.data.Component /
AtomProxy (fr3d/data/components.py:118-202, fr3d/data/base.py:80-209) into the
per-nt device record documented in include/fr3d_b200.h.  Two sources:

  pack_specs(specs)            plain coordinate specs (what a CIF atom_site loop gives);
                               frames and hydrogens are then produced on the GPU (K1)
  pack_structures(structures)  duck-typed fr3d Structure/Component objects that already
                               carry rotation_matrix / centers / inferred hydrogens
"""

import operator

import numpy as np

from . import _lib
from . import definitions as defs

_ATOM_FIELDS = operator.attrgetter("name", "x", "y", "z")        # one C-level call per atom object
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
                sof = self.structure_of().tolist()
                pdb, model, chain, seq, num = self.pdb, self.model, self.chain, self.sequence, self.number
                alt, ins, sym = self.alt_id, self.insertion_code, self.symmetry
                out = []
                for k in range(self.n):
                    nk = num[k]
                    # the common id: no alt id, no insertion code, identity operator, a residue number - nothing is
                    # trimmed after the number, so the id is the first five fields joined (encode_unit_id)
                    if not alt[k] and not ins[k] and (not sym[k] or sym[k] == '1_555') and nk is not None and nk != '':
                        p_, m_, c_, s_ = pdb[sof[k]], model[k], chain[k], seq[k]
                        out.append("%s|%s|%s|%s|%s" % ('' if p_ is None else p_, '' if m_ is None else m_,
                                                       '' if c_ is None else c_, '' if s_ is None else s_, nk))
                    else:
                        out.append(encode_unit_id(pdb[sof[k]], model[k], chain[k], seq[k], nk, alt[k], ins[k], sym[k]))
                self._unit_ids = out
        return self._unit_ids

    def unit_ids_of(self, s):
        """Unit ids of structure s only (list; position = nt - struct_off[s])."""
        lo, hi = int(self.struct_off[s]), int(self.struct_off[s + 1])
        if self._unit_ids is not None:
            return self._unit_ids[lo:hi]
        if self._unit_id_fn is not None:
            return self.unit_ids()[lo:hi]
        pdb, model, chain, seq, num = self.pdb[s], self.model, self.chain, self.sequence, self.number
        alt, ins, sym = self.alt_id, self.insertion_code, self.symmetry
        p_ = '' if pdb is None else pdb
        out = []
        for k in range(lo, hi):
            nk = num[k]
            if not alt[k] and not ins[k] and (not sym[k] or sym[k] == '1_555') and nk is not None and nk != '':
                m_, c_, s_ = model[k], chain[k], seq[k]
                out.append("%s|%s|%s|%s|%s" % (p_, '' if m_ is None else m_, '' if c_ is None else c_,
                                               '' if s_ is None else s_, nk))
            else:
                out.append(encode_unit_id(pdb, model[k], chain[k], seq[k], nk, alt[k], ins[k], sym[k]))
        return out

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
    if parent is None or parent not in defs.PARENT_CODE:
        # unknown residue, or a modified residue whose parent is DA / DC / DG: the reference gives those the parent
        # "DA" ... (NA:1257), for which it has planes but no cutoff boxes; they are left without a frame here and drop
        # out of the annotation instead of being paired as their RNA counterparts
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


_ZERO3 = [0.0, 0.0, 0.0]
_ZERO_BASE = [0.0] * (3 * _lib.BASE_SLOTS)
_ZERO_BB = [0.0] * (3 * _lib.BB_SLOTS)


class _Rows(object):
    """Per-nt records collected as plain Python lists while the atoms are walked, written to the numpy
    arrays of the batch in one go (element-wise numpy writes dominated the packing time)."""

    def __init__(self):
        self.base, self.bb, self.bmask, self.bbmask = [], [], [], []
        self.meta0, self.meta1, self.meta7 = [], [], []

    def add_unknown(self):
        self.base += _ZERO_BASE; self.bb += _ZERO_BB
        self.bmask.append(0); self.bbmask.append(0)
        self.meta0.append(-1); self.meta1.append(0); self.meta7.append(0)

    def flush(self, batch):
        if not batch.n:
            return
        batch.base_xyz[:] = np.array(self.base, dtype=np.float64).reshape(batch.n, _lib.BASE_SLOTS, 3)   # flat lists:
        batch.bb_xyz[:] = np.array(self.bb, dtype=np.float64).reshape(batch.n, _lib.BB_SLOTS, 3)         # nested ones parse slowly
        batch.base_mask[:] = np.array(self.bmask, dtype=np.uint32)
        batch.bb_mask[:] = np.array(self.bbmask, dtype=np.uint32)
        batch.meta[:, 0] = self.meta0
        batch.meta[:, 1] = self.meta1
        batch.meta[:, 7] = self.meta7


def _fill_atoms(rows, maps, atoms, frames_given=False):
    """Append one nt to rows.  atoms: iterable of (name, x, y, z).  Duplicate names are averaged like AtomProxy
    (fr3d/data/base.py:150-166).  Modified residues: see FR3D_FLAG_DUP in include/fr3d_b200.h."""
    pcode, flags, bslot, bbslot, dupmask = maps
    base = [None] * _lib.BASE_SLOTS          # slot -> [x, y, z, count]
    bb = [None] * _lib.BB_SLOTS
    saw_h = False
    for name, x, y, z in atoms:
        if 'H' in name:
            saw_h = True
        s = bslot.get(name)
        if s is not None:
            tgt = base
        else:
            s = bbslot.get(name)
            if s is None:
                continue
            tgt = bb
        cur = tgt[s]
        if cur is None:
            tgt[s] = [x, y, z, 1]
        else:
            c = cur[3]
            cur[0] = (cur[0] * c + x) / (c + 1)
            cur[1] = (cur[1] * c + y) / (c + 1)
            cur[2] = (cur[2] * c + z) / (c + 1)
            cur[3] = c + 1
    bmask = 0
    bbmask = 0
    for s_, cur in enumerate(base):
        if cur is not None:
            bmask |= 1 << s_
    for s_, cur in enumerate(bb):
        if cur is not None:
            bbmask |= 1 << s_
    meta7 = 0
    if flags & _lib.FLAG_MODIFIED:
        if frames_given:
            # Component objects of the reference already hold the ideal copies: twin slots are those seen twice
            twin = 0
            for s_, cur in enumerate(base):
                if cur is not None and cur[3] == 2:
                    twin |= 1 << s_
            if twin:
                flags |= _lib.FLAG_DUP
                meta7 = twin << 16
        elif not saw_h:
            flags |= _lib.FLAG_DUP
            meta7 = dupmask
    flat = rows.base
    for cur in base:
        flat += cur[:3] if cur is not None else _ZERO3
    flat = rows.bb
    for cur in bb:
        flat += cur[:3] if cur is not None else _ZERO3
    rows.bmask.append(bmask | ((pcode + 1) & 15) << 16 | (flags & 15) << 20)      # tag read by the classify kernel
    rows.bbmask.append(bbmask)
    rows.meta0.append(pcode); rows.meta1.append(flags); rows.meta7.append(meta7)


def _fix_amino_hydrogens(rows, parent):
    """components.py:405-445, on the nt just appended to rows: observed amino hydrogens are relabelled so that
    H62 / H42 / H22 is the one nearer N7 / C5 / N1."""
    if parent not in _AMINO:
        return
    heavy, h1, h2 = _AMINO[parent]
    so = defs.SLOT_OF[parent]
    sh, s1, s2 = so[heavy], so[h1], so[h2]
    m = rows.bmask[-1]
    if (m >> sh & 1) and (m >> s1 & 1) and (m >> s2 & 1):
        flat = rows.base
        o = len(flat) - 3 * _lib.BASE_SLOTS                  # the nt just appended
        a, p1, p2 = flat[o + 3 * sh:o + 3 * sh + 3], flat[o + 3 * s1:o + 3 * s1 + 3], flat[o + 3 * s2:o + 3 * s2 + 3]
        d1 = (a[0] - p1[0]) ** 2 + (a[1] - p1[1]) ** 2 + (a[2] - p1[2]) ** 2
        d2 = (a[0] - p2[0]) ** 2 + (a[1] - p2[1]) ** 2 + (a[2] - p2[2]) ** 2
        if d1 < d2:
            flat[o + 3 * s1:o + 3 * s1 + 3] = p2
            flat[o + 3 * s2:o + 3 * s2 + 3] = p1


def _prev_o3_of_structure(lo, hi, model, symmetry, chain, m4, uid, has_o3, src, dst):
    """map_unit_id_to_previous_O3 (NA:480-525) for one structure, any mix of models / operators / repeated positions:
    walk the nts in (model, symmetry, chain, index, unit id) order, link successive indices of one chain."""
    order = sorted(range(lo, hi), key=lambda k: (model[k], symmetry[k] or "", chain[k], m4[k], uid[k]))
    prev = None
    for k in order:
        if prev is not None and model[k] == model[prev] and symmetry[k] == symmetry[prev] \
                and chain[k] == chain[prev] and m4[k] == m4[prev] + 1:
            if has_o3[prev]:                     # O3' of the previous nt
                src.append(prev); dst.append(k)
        prev = k


def _finish_identity(batch):
    """group ids, chain ranks, alt-id keys and previous-O3' (map_unit_id_to_previous_O3, NA:480-525)."""
    n = batch.n
    if not n:
        return
    sof = batch.structure_of()
    sof_l = sof.tolist()
    chain_rank = {c: r for r, c in enumerate(sorted(set(batch.chain)))}
    groups, resid, alts = {}, {}, {}
    # one pass per column (list comprehensions), one numpy write per column
    m2 = [groups.setdefault(key, len(groups)) for key in zip(sof_l, batch.model)]
    m3 = [chain_rank[c] for c in batch.chain]
    m4 = [0 if i is None else int(i) for i in batch.index]
    if any(a is not None for a in batch.alt_id):     # alt-id twins screen, NA:706-713
        m5 = [resid.setdefault(key, len(resid)) for key in zip(sof_l, batch.number, batch.sequence, batch.chain,
                                                                batch.symmetry, batch.insertion_code)]
        m6 = [alts.setdefault(a, len(alts)) for a in batch.alt_id]
    else:
        m5, m6 = np.arange(n), 0
    batch.meta[:, 2] = m2; batch.meta[:, 3] = m3; batch.meta[:, 4] = m4; batch.meta[:, 5] = m5; batch.meta[:, 6] = m6
    # previous O3'.  Structures with one model, one operator and no repeated (chain, index) - nearly all - are done
    # together with numpy: sort by (structure, chain rank, index), link neighbours; the others one by one.
    has_o3 = (batch.bb_mask >> 5) & 1
    m3a, m4a = np.asarray(m3, dtype=np.int64), np.asarray(m4, dtype=np.int64)
    codes = {}
    ms = np.fromiter((codes.setdefault(key, len(codes)) for key in zip(batch.model, batch.symmetry)), dtype=np.int64, count=n)
    starts = batch.struct_off[:-1][np.diff(batch.struct_off) > 0]
    simple = np.zeros(batch.n_structures, dtype=bool)
    nonempty = np.nonzero(np.diff(batch.struct_off) > 0)[0]
    simple[nonempty] = np.minimum.reduceat(ms, starts) == np.maximum.reduceat(ms, starts)
    order = np.lexsort((m4a, m3a, sof))
    so, c3, c4 = sof[order], m3a[order], m4a[order]
    tie = (so[1:] == so[:-1]) & (c3[1:] == c3[:-1]) & (c4[1:] == c4[:-1])
    simple[so[1:][tie]] = False
    link = (so[1:] == so[:-1]) & (c3[1:] == c3[:-1]) & (c4[1:] == c4[:-1] + 1) & simple[so[1:]] & (has_o3[order[:-1]] == 1)
    src, dst = order[:-1][link].tolist(), order[1:][link].tolist()
    if not simple.all():
        uid = batch.unit_ids()
        has_l = has_o3.tolist()
        for s in np.nonzero(~simple)[0].tolist():
            _prev_o3_of_structure(int(batch.struct_off[s]), int(batch.struct_off[s + 1]), batch.model, batch.symmetry,
                                  batch.chain, m4, uid, has_l, src, dst)
    if dst:
        batch.bb_xyz[dst, defs.BB_PREV_O3] = batch.bb_xyz[src, 5]
        batch.bb_mask[dst] |= np.uint32(1 << defs.BB_PREV_O3)


def pack_specs(specs):
    """specs: list of {"pdb", "nts": [{"chain","sequence","number","index","model","symmetry",
    "alt_id","insertion_code","atoms":[[name,x,y,z],...]}]}.  Only RNA/DNA residues belong in a spec."""
    if isinstance(specs, dict):
        specs = [specs]
    n = sum(len(s["nts"]) for s in specs)
    b = NtBatch(n)
    rows = _Rows()
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
                rows.add_unknown()
            else:
                _fill_atoms(rows, maps, nt["atoms"])
                if not (maps[1] & _lib.FLAG_MODIFIED):
                    _fix_amino_hydrogens(rows, defs.PARENTS[maps[0]])
            k += 1
        off.append(k)
    rows.flush(b)
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
    rows = _Rows()
    off = [0]
    k = 0
    uids = []
    framed, rots, ctrs = [], [], []
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
                rows.add_unknown()
            else:
                _fill_atoms(rows, maps, map(_ATOM_FIELDS, c.atoms()), frames_given=True)
                R = c.rotation_matrix
                ctr = c.centers["base"]
                if R is not None and len(ctr) == 3:
                    framed.append(k)
                    rots.append(np.asarray(R, dtype=float).reshape(9))
                    ctrs.append(np.asarray(ctr, dtype=float))
                    rows.bmask[-1] |= _lib.MASK_HAS_FRAME
            k += 1
        off.append(k)
    rows.flush(b)
    if framed:
        b.rot[framed] = np.array(rots)
        b.center[framed] = np.array(ctrs)
    b.struct_off = np.array(off, dtype=np.int64)
    b._unit_ids = uids
    _finish_identity(b)
    return b


# ------------------------------------------------------------------ array-native ingest
def specs_to_columns(specs):
    """Column form of coordinate specs: what a columnar mmCIF atom_site reader yields after residues have been
    delimited (the reference does that in fr3d/cif/reader.py:440-481).  Returns the arguments of pack_columns."""
    if isinstance(specs, dict):
        specs = [specs]
    res = {k: [] for k in ("sequence", "chain", "number", "index", "model", "symmetry", "alt_id", "insertion_code")}
    atom_res, atom_name, atom_xyz, off, pdb = [], [], [], [0], []
    r = 0
    for spec in specs:
        pdb.append(spec["pdb"])
        for nt in spec["nts"]:
            for k in res:
                res[k].append(nt.get(k))
            for a in nt["atoms"]:
                atom_res.append(r); atom_name.append(a[0]); atom_xyz.append(a[1:4])
            r += 1
        off.append(r)
    return (np.array(atom_res, dtype=np.int64), np.array(atom_name, dtype=str),
            np.array(atom_xyz, dtype=np.float64).reshape(-1, 3), res, pdb, np.array(off, dtype=np.int64))


def pack_columns(atom_res, atom_name, atom_xyz, res, pdb, struct_off):
    """pack_specs for column data, without a Python loop over atoms.

    atom_res int[n_atoms]: residue (row of the residue table) of every atom, atom_name str[n_atoms], atom_xyz
    float64[n_atoms, 3]; res: dict of per-residue columns (sequence, chain, number, index, model, symmetry, alt_id,
    insertion_code); pdb: id per structure; struct_off: residue offsets of the structures.  Produces exactly the
    NtBatch pack_specs produces (tests/test_host_logic.py compares every array): atom names are mapped to slots
    through one (residue type x atom name) table, coordinates are scattered with one fancy assignment, masks with
    bitwise-or reductions.  Residues in which a name occurs more than once (alternate conformers listed together;
    AtomProxy averages them, base.py:150-166) take the per-atom path of pack_specs."""
    atom_res = np.asarray(atom_res, dtype=np.int64)
    atom_xyz = np.asarray(atom_xyz, dtype=np.float64).reshape(-1, 3)
    names = np.asarray(atom_name)
    if names.dtype.kind not in "US":
        names = names.astype(str)
    seqs = list(res["sequence"])
    n = len(seqs)
    b = NtBatch(n)
    b.pdb = list(pdb)
    b.struct_off = np.asarray(struct_off, dtype=np.int64)
    b.model, b.chain, b.sequence = list(res["model"]), list(res["chain"]), seqs
    b.number, b.index, b.symmetry = list(res["number"]), list(res["index"]), list(res["symmetry"])
    b.alt_id = list(res.get("alt_id") or [None] * n)
    b.insertion_code = list(res.get("insertion_code") or [None] * n)
    if n == 0:
        _finish_identity(b)
        return b
    # residue types and atom names as small integer codes
    useq = sorted(set(seqs))
    seq_code = {s: k for k, s in enumerate(useq)}
    sc_res = np.fromiter((seq_code[s] for s in seqs), dtype=np.int64, count=n)
    if len(names) == 0:
        uname, nc = np.zeros(0, dtype=str), np.zeros(0, dtype=np.int64)
    elif names.dtype.kind == "U" and names.dtype.itemsize <= 16:
        # names of at most four characters (every PDB atom name): pack the code points into one uint64 per atom and
        # find the distinct names on integers - sorting 7 k strings per structure was the largest single cost
        cp = np.ascontiguousarray(names).view(np.uint32).reshape(len(names), -1).astype(np.uint64)
        if cp.max() < 65536:
            key = cp[:, 0].copy()
            for c in range(1, cp.shape[1]):
                key |= cp[:, c] << np.uint64(16 * c)
            ukey, first, nc = np.unique(key, return_index=True, return_inverse=True)
            uname = names[first].astype(str)
        else:
            uname, nc = np.unique(names, return_inverse=True)
            uname = uname.astype(str)
    else:
        uname, nc = np.unique(names, return_inverse=True)
        uname = uname.astype(str)
    maps = [_slot_maps(s) for s in useq]
    t_base = np.full((len(useq), max(len(uname), 1)), -1, dtype=np.int64)
    t_bb = np.full((len(useq), max(len(uname), 1)), -1, dtype=np.int64)
    name_code = {nm: k for k, nm in enumerate(uname.tolist())}
    for k, m in enumerate(maps):
        if m is None:
            continue
        for nm, slot in m[2].items():
            if nm in name_code:
                t_base[k, name_code[nm]] = slot
        for nm, slot in m[3].items():
            if nm in name_code and t_base[k, name_code[nm]] < 0:          # (_fill_atoms looks the base map up first)
                t_bb[k, name_code[nm]] = slot
    pcode = np.array([-1 if m is None else m[0] for m in maps], dtype=np.int64)[sc_res]
    flags = np.array([0 if m is None else m[1] for m in maps], dtype=np.int64)[sc_res]
    dupmask = np.array([0 if m is None else m[4] for m in maps], dtype=np.int64)[sc_res]
    sc_atom = sc_res[atom_res]
    sb = t_base[sc_atom, nc] if len(nc) else np.zeros(0, dtype=np.int64)
    sk = t_bb[sc_atom, nc] if len(nc) else np.zeros(0, dtype=np.int64)
    in_base, in_bb = sb >= 0, sk >= 0
    # residues with a repeated (residue, slot): per-atom path afterwards
    key = np.concatenate([atom_res[in_base] * 32 + sb[in_base], atom_res[in_bb] * 32 + 16 + sk[in_bb]])
    uk, cnt = np.unique(key, return_counts=True)
    slow = np.unique(uk[cnt > 1] // 32)
    is_slow = np.zeros(n, dtype=bool)
    is_slow[slow] = True
    fast_b = in_base & ~is_slow[atom_res]
    fast_k = in_bb & ~is_slow[atom_res]
    b.base_xyz[atom_res[fast_b], sb[fast_b]] = atom_xyz[fast_b]
    b.bb_xyz[atom_res[fast_k], sk[fast_k]] = atom_xyz[fast_k]
    bmask = np.zeros(n, dtype=np.int64)
    bbmask = np.zeros(n, dtype=np.int64)
    np.bitwise_or.at(bmask, atom_res[fast_b], np.left_shift(1, sb[fast_b]))
    np.bitwise_or.at(bbmask, atom_res[fast_k], np.left_shift(1, sk[fast_k]))
    # modified residues without any observed hydrogen receive ideal-position copies (components.py:505-513)
    has_h_name = np.array(['H' in nm for nm in uname.tolist()], dtype=bool)
    saw_h = np.zeros(n, dtype=bool)
    if len(nc):
        np.logical_or.at(saw_h, atom_res, has_h_name[nc])
    dup = ((flags & _lib.FLAG_MODIFIED) != 0) & ~saw_h & (pcode >= 0)
    flags = np.where(dup, flags | _lib.FLAG_DUP, flags)
    meta7 = np.where(dup, dupmask, 0)
    # observed amino hydrogens: H62 / H42 / H22 is the one nearer N7 / C5 / N1 (components.py:405-445)
    for parent, (heavy, h1, h2) in _AMINO.items():
        so = defs.SLOT_OF[parent]
        sh, s1, s2 = so[heavy], so[h1], so[h2]
        need = (1 << sh) | (1 << s1) | (1 << s2)
        rows = np.nonzero((pcode == defs.PARENT_CODE[parent]) & ((flags & _lib.FLAG_MODIFIED) == 0) &
                          ((bmask & need) == need) & ~is_slow)[0]
        if len(rows):
            a, p1, p2 = b.base_xyz[rows, sh], b.base_xyz[rows, s1].copy(), b.base_xyz[rows, s2].copy()
            d1 = ((a - p1) ** 2).sum(axis=1)        # same order of operations as _fix_amino_hydrogens
            d2 = ((a - p2) ** 2).sum(axis=1)
            swap = rows[d1 < d2]
            b.base_xyz[swap, s1] = p2[d1 < d2]
            b.base_xyz[swap, s2] = p1[d1 < d2]
    known = pcode >= 0
    b.base_mask[:] = np.where(known, bmask | (((pcode + 1) & 15) << 16) | ((flags & 15) << 20), 0).astype(np.uint32)
    b.bb_mask[:] = np.where(known, bbmask, 0).astype(np.uint32)
    b.meta[:, 0] = pcode
    b.meta[:, 1] = np.where(known, flags, 0)
    b.meta[:, 7] = np.where(known, meta7, 0)
    if len(slow):                                    # the rare residues with repeated atom names
        order = np.argsort(atom_res, kind="stable")
        starts = np.searchsorted(atom_res[order], np.arange(n + 1))
        for r_ in slow.tolist():
            m = maps[sc_res[r_]]
            if m is None:
                continue
            rows_ = _Rows()
            idx = order[starts[r_]:starts[r_ + 1]]
            _fill_atoms(rows_, m, [(str(names[i]), atom_xyz[i, 0], atom_xyz[i, 1], atom_xyz[i, 2]) for i in idx.tolist()])
            if not (m[1] & _lib.FLAG_MODIFIED):
                _fix_amino_hydrogens(rows_, defs.PARENTS[m[0]])
            b.base_xyz[r_] = np.array(rows_.base, dtype=np.float64).reshape(_lib.BASE_SLOTS, 3)
            b.bb_xyz[r_] = np.array(rows_.bb, dtype=np.float64).reshape(_lib.BB_SLOTS, 3)
            b.base_mask[r_] = rows_.bmask[0]
            b.bb_mask[r_] = rows_.bbmask[0]
            b.meta[r_, 0], b.meta[r_, 1], b.meta[r_, 7] = rows_.meta0[0], rows_.meta1[0], rows_.meta7[0]
    _finish_identity(b)
    return b
