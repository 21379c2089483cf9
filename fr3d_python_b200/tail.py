"""Host side of the device tail (csrc/tail.cuh, fr3d_annotation_tail): label-id tables, the identity arrays the
kernel reads, and the array-native result with lazily built reference objects.

The device works on integers only.  This module owns the id space:

  LabelTable     every interaction label the reference can append for a given cutoff-box table, as a list of
                 strings; per label the ids of reverse_edges(label) and "n" + label and the flags the tail
                 needs (NA:1171-1177 duplication, "n" prefix, nested-cWW family NA:1077, edge letter NA:543);
                 per cutoff box the label ids recorded for u1 / u2 and the hydrogen-bond atom sets of
                 store_basepair_quality (NA:2345-2361) as bit masks.
  TailIdent      per nt: chain-table entry (structure, chain id) and the rank of "model symmetry chain"
                 (NA:1192); per chain: the length of its nested-cWW endpoint array (NA:1106).
  BatchAnnotations   rows (label id, nt1, nt2, crossing) per structure in the reference's append order;
                 result[s] builds the reference's (interaction_to_list_of_tuples, category_to_interactions)
                 for structure s on first access.
"""

from collections import defaultdict

import numpy as np

from . import _lib
from .postprocess import reverse_edges, _DUP_STACK, _NESTED_FAMILIES, _SR, so_labels, stack_labels  # noqa: F401

ROW_DTYPE = np.dtype([('label', '<i4'), ('nt1', '<i4'), ('nt2', '<i4'), ('crossing', '<i4')])
TAIL_LABEL_DTYPE = np.dtype([('rev', '<i4'), ('near', '<i4'), ('flags', '<u4'), ('pad', '<i4')])
TAIL_BOX_DTYPE = np.dtype([('label', '<i4', (2, 2)), ('atoms1', '<u8'), ('atoms2', '<u8')])
assert ROW_DTYPE.itemsize == 16 and TAIL_LABEL_DTYPE.itemsize == 16 and TAIL_BOX_DTYPE.itemsize == 32


def _duplicated(label):
    """NA:1171-1177: rows of these labels are appended a second time under reverse_edges(label)."""
    return label in _DUP_STACK or label[0] in "ct" or label[0:2] in ("nc", "nt")


class LabelTable(object):
    def __init__(self, boxtable, box_atoms):
        self.strings = []
        self.ids = {}
        self.category = []            # per label: (category, labels added to category_to_interactions) or None
        depth = []                    # 0: a box name or "n" + name (or its reverse); 1: "n" + that; -1: not a basepair

        def lid(s, d, cat):
            k = self.ids.get(s)
            if k is None:
                k = len(self.strings)
                self.ids[s] = k
                self.strings.append(s)
                depth.append(d)
                self.category.append(cat)
            return k

        slot_label = np.full((8, 32, 2), -1, dtype=np.int32)
        for code in range(32):
            if (code & 7) < 6:                                           # NA:1386-1459
                l, r = so_labels(code)
                pair = (lid(l, -1, ('sO', (l,))), lid(r, -1, ('sO', (r,))))
                slot_label[_lib.SLOT_SO12, code] = pair
                slot_label[_lib.SLOT_SO21, code] = pair
            if code < 8:                                                 # NA:1682-1717
                l, r = stack_labels(code)
                slot_label[_lib.SLOT_STACK, code, 0] = lid(l, -1, ('stacking', (l, r)))
                lid(r, -1, ('stacking', (r, l)))
            if code < 2:
                slot_label[_lib.SLOT_SR12, code, 0] = slot_label[_lib.SLOT_SR21, code, 0] = \
                    lid(_SR[code], -1, ('sugar_ribose', (_SR[code],)))
            if (code & 15) < 10:                                         # NA:1973-2019
                pre = "n" if code & 16 else ""
                l = pre + str(code & 15) + "BPh"
                slot_label[_lib.SLOT_BPH, code, 0] = lid(l, -1, ('backbone', (l,)))
                l = pre + str(code & 15) + "BR"
                slot_label[_lib.SLOT_BR, code, 0] = lid(l, -1, ('backbone', (l,)))
            slot_label[_lib.SLOT_CP, code, 0] = lid('cp', -1, ('coplanar', ('cp',)))

        def bp(s, d):
            return lid(s, d, ('basepair', (s, reverse_edges(s))))

        atom_bit = {}
        boxes = np.zeros(max(len(boxtable.box_name), 1), dtype=TAIL_BOX_DTYPE)
        for b, name in enumerate(boxtable.box_name):
            for near in (0, 1):
                lw = ("n" + name) if near else name                      # NA:2693-2696
                boxes[b]['label'][0][near] = bp(lw, 0)
                boxes[b]['label'][1][near] = bp(reverse_edges(lw), 0)    # NA:1013
            for field, names in (('atoms1', box_atoms[b][0]), ('atoms2', box_atoms[b][1])):
                m = 0
                for a in names:
                    if a not in atom_bit:
                        if len(atom_bit) >= 64:
                            raise _lib.Fr3dError("more than 64 distinct hydrogen-bond atom names")
                        atom_bit[a] = len(atom_bit)
                    m |= 1 << atom_bit[a]
                boxes[b][field] = m
        self.hydrogen_atoms = sum(1 << k for a, k in atom_bit.items() if "H" in a)      # NA:581-582
        # closure: reverse_edges of every duplicated label, and "n" + label for the labels an entry can carry (the
        # emission demotes an entry at most once, NA:1031-1032)
        rev, near = [], []
        k = 0
        while k < len(self.strings):
            s = self.strings[k]
            d = depth[k]
            if _duplicated(s):
                r = reverse_edges(s)
                rev.append(bp(r, d) if d >= 0 else lid(r, -1, None))
            else:
                rev.append(-1)
            near.append(bp("n" + s, 1) if d == 0 else -1)
            k += 1
        if len(self.strings) > _lib.TAIL_MAX_LABELS:
            raise _lib.Fr3dError("%d interaction labels, the device tail supports %d" % (len(self.strings), _lib.TAIL_MAX_LABELS))
        labels = np.zeros(len(self.strings), dtype=TAIL_LABEL_DTYPE)
        for k, s in enumerate(self.strings):
            clean = s.replace("n", "")
            edge = ord(clean[1].lower()) & 0xFF if len(clean) > 1 else 0          # NA:543
            labels[k] = (rev[k], near[k], (1 if _duplicated(s) else 0) | (2 if s.startswith("n") else 0) |
                         (4 if s in _NESTED_FAMILIES else 0) | (edge << 8), 0)
        self.labels, self.boxes, self.slot_label = labels, boxes, slot_label
        self._dev = {}

    def name(self, label_id):
        if label_id & _lib.LABEL_COVALENT:
            return "p_%d" % (label_id & (_lib.LABEL_COVALENT - 1))
        return self.strings[label_id]

    def device(self, torch, device):
        """(TailTablesStruct, tensors that keep the device copies alive)."""
        key = str(device)
        if key not in self._dev:
            parts = [self.labels.view(np.uint8).reshape(-1), self.boxes.view(np.uint8).reshape(-1),
                     self.slot_label.view(np.uint8).reshape(-1)]
            offs, total = [], 0
            for p in parts:
                offs.append(total)
                total += (len(p) + 255) // 256 * 256
            blob = np.zeros(total, dtype=np.uint8)
            for o, p in zip(offs, parts):
                blob[o:o + len(p)] = p
            t = torch.from_numpy(blob).to(device)
            base = t.data_ptr()
            st = _lib.TailTablesStruct(n_labels=len(self.labels), n_boxes=len(self.boxes), labels=base + offs[0],
                                       boxes=base + offs[1], slot_label=base + offs[2],
                                       hydrogen_atoms=self.hydrogen_atoms)
            self._dev[key] = (st, t)
        return self._dev[key]


class TailIdent(object):
    """Identity arrays of a batch for the device tail (host numpy; see fr3d_tail_ident_t)."""

    def __init__(self, batch):
        n, S = batch.n, batch.n_structures
        self.n, self.n_struct = n, S
        self.struct_off = np.asarray(batch.struct_off, dtype=np.int64)
        sizes = np.diff(self.struct_off)
        if S and sizes.max() > _lib.TAIL_MAX_NT:
            raise _lib.Fr3dError("a structure of %d nts: the device tail supports %d per structure (tail='host' "
                                 "has no limit)" % (int(sizes.max()), _lib.TAIL_MAX_NT))
        index = batch.meta[:, 4].astype(np.int64)
        if n and (index.min() < 0 or index.max() >= _lib.TAIL_MAX_INDEX):
            raise _lib.Fr3dError("nt index outside [0, 2^23): not supported by the device tail (tail='host' is)")
        crank = batch.meta[:, 3].astype(np.int64)
        if n:
            # chain table = distinct (structure, chain id), found on the runs of equal keys (nts of one chain are
            # contiguous but for models / symmetry copies, so there are a few runs per structure, not one per nt)
            sof = batch.structure_of().astype(np.int64)
            K = int(crank.max()) + 1
            key = sof * K + crank
            start = np.concatenate([[0], np.nonzero(key[1:] != key[:-1])[0] + 1])
            run_key = key[start]
            run_max = np.maximum.reduceat(index, start)
            uniq, run_inv = np.unique(run_key, return_inverse=True)
            mx = np.zeros(len(uniq), dtype=np.int64)
            np.maximum.at(mx, run_inv, run_max)
            self.nt_chain = np.repeat(run_inv, np.diff(np.concatenate([start, [n]]))).astype(np.int32)
            self.struct_chain_off = np.searchsorted(uniq // K, np.arange(S + 1)).astype(np.int32)
            ends = np.concatenate([[0], np.cumsum(mx + 1)])              # chain_to_max_index + 1 (NA:1072, 1106)
        else:
            self.nt_chain = np.zeros(0, dtype=np.int32)
            self.struct_chain_off = np.zeros(S + 1, dtype=np.int32)
            ends = np.zeros(1, dtype=np.int64)
        if ends[-1] >= 2 ** 31:
            raise _lib.Fr3dError("nested-cWW endpoint arrays need %d entries" % int(ends[-1]))
        self.chain_ends_off = ends.astype(np.int32)
        self.n_chains = len(ends) - 1
        self.n_ends = int(ends[-1])
        self.nt_cov = self._covalent_ranks(batch, crank)

    @staticmethod
    def _covalent_ranks(batch, crank):
        """Rank of nt.model + " " + nt.symmetry + " " + nt.chain (NA:1192) among the batch's strings.  With one
        model and one operator in the whole batch the strings differ in the chain id only and sort like it."""
        n = batch.n
        if n == 0:
            return np.zeros(0, dtype=np.int32)
        if len(set(batch.model)) == 1 and len(set(batch.symmetry)) == 1:
            return crank.astype(np.int32)
        codes = {}
        ids = np.fromiter((codes.setdefault(key, len(codes)) for key in zip(batch.model, batch.symmetry, batch.chain)),
                          dtype=np.int64, count=n)
        names = ["%s %s %s" % ("" if m is None else m, "" if s is None else s, "" if c is None else c)
                 for (m, s, c) in codes.keys()]
        order = sorted(range(len(names)), key=lambda k: names[k])
        rank_of_name = {}
        r = -1
        prev = None
        for k in order:                       # equal strings (e.g. model 1 vs "1") share a rank
            if names[k] != prev:
                r += 1
                prev = names[k]
            rank_of_name[k] = r
        if r >= (1 << 23):
            raise _lib.Fr3dError("more than 2^23 distinct model / symmetry / chain strings")
        return np.array([rank_of_name[k] for k in range(len(names))], dtype=np.int32)[ids]

    def chunk(self, s0, s1):
        """Arrays of structures s0..s1-1 with offsets local to the chunk: (struct_off, struct_chain_off,
        chain_ends_off, nt_chain, nt_cov, n_chains, n_ends)."""
        lo, hi = int(self.struct_off[s0]), int(self.struct_off[s1])
        c0, c1 = int(self.struct_chain_off[s0]), int(self.struct_chain_off[s1])
        e0 = int(self.chain_ends_off[c0])
        return ((self.struct_off[s0:s1 + 1] - lo).astype(np.int32),
                (self.struct_chain_off[s0:s1 + 1] - c0).astype(np.int32),
                (self.chain_ends_off[c0:c1 + 1] - e0).astype(np.int32),
                (self.nt_chain[lo:hi] - c0).astype(np.int32), self.nt_cov[lo:hi],
                c1 - c0, int(self.chain_ends_off[c1]) - e0)


def tail_ident(batch):
    ti = getattr(batch, "_tail_ident", None)
    if ti is None or ti.n != batch.n:
        ti = TailIdent(batch)
        batch._tail_ident = ti
    return ti


class BatchAnnotations(object):
    """Result of annotate_batch: per-structure rows in the reference's append order.

    len(result) structures; result[s] -> (interaction_to_list_of_tuples, category_to_interactions) exactly like
    annotate_nt_nt_in_structure returns them (built on first access, cached); result.rows_of(s) -> the structured
    row array (label id, nt1, nt2, crossing) with GLOBAL nt indices of the batch; result.labels.name(id) -> str."""

    def __init__(self, batch, rows, row_start, row_count, labels, n_pairs=None, nt_base=None):
        """rows[row_start[s] : row_start[s] + row_count[s]] are the rows of structure s; their nt fields count from
        nt_base[s] for the first nt of the structure (default batch.struct_off: global indices of the batch; a
        sharded run delivers shard-local ones)."""
        self.batch, self.rows, self.row_start, self.row_count, self.labels = batch, rows, row_start, row_count, labels
        self.n_pairs = n_pairs
        self.nt_base = nt_base
        self._cache = {}

    def __len__(self):
        return len(self.row_start)

    def rows_of(self, s):
        """Rows of structure s with GLOBAL nt indices of the batch (a view when the stored rows already are)."""
        a = int(self.row_start[s])
        r = self.rows[a:a + int(self.row_count[s])]
        if self.nt_base is not None:
            shift = int(self.batch.struct_off[s]) - int(self.nt_base[s])
            if shift:
                r = r.copy()
                r['nt1'] += shift
                r['nt2'] += shift
        return r

    @property
    def n_rows(self):
        return int(np.sum(self.row_count))

    def __iter__(self):
        for s in range(len(self)):
            yield self[s]

    def __getitem__(self, s):
        if isinstance(s, slice):
            return [self[k] for k in range(*s.indices(len(self)))]
        if s < 0:
            s += len(self)
        res = self._cache.get(s)
        if res is None:
            res = self._build(s)
            self._cache[s] = res
        return res

    def _build(self, s):
        """The reference's (interaction_to_list_of_tuples, category_to_interactions) of structure s.  The rows arrive grouped
        by label (append order = label by label), so every run of one label becomes its list with C-level map / zip:
        one tuple allocation per row and no Python bytecode per row."""
        r = self.rows_of(s)
        lo = int(self.batch.struct_off[s])
        uid = self.batch.unit_ids_of(s)
        i2t = defaultdict(list)
        c2i = defaultdict(set)
        if len(r) == 0:
            return i2t, c2i
        cats = self.labels.category
        lab = r['label']
        cuts = np.flatnonzero(lab[1:] != lab[:-1]) + 1
        starts = [0] + cuts.tolist()
        ends = cuts.tolist() + [len(r)]
        get = uid.__getitem__
        u1 = list(map(get, (r['nt1'] - lo).tolist()))
        u2 = list(map(get, (r['nt2'] - lo).tolist()))
        cr = r['crossing'].astype(object)
        cr[r['crossing'] < 0] = None
        cr = cr.tolist()
        seen = set()
        for b, e, l in zip(starts, ends, lab[starts].tolist()):
            nm = self.labels.name(l)
            if l not in seen:
                seen.add(l)
                if l & _lib.LABEL_COVALENT:
                    c2i['covalent'].add(nm)
                elif cats[l] is not None:
                    cat, adds = cats[l]
                    c2i[cat].update(adds)
                    if cat == 'basepair':
                        c2i['basepair_detail'].update(adds)
            i2t[nm].extend(zip(u1[b:e], u2[b:e], cr[b:e]))
        return i2t, c2i
