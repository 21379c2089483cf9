"""Device side of the annotator: owns torch buffers, calls the C ABI (K1-K4).

PyTorch is used only for device memory, pinned host memory and streams.
Everything fails loudly when CUDA or the extension is unavailable — there is
no CPU path in this package.
"""

import ctypes as C
import os

import numpy as np

from . import _lib
from . import definitions as defs

HIT_DTYPE = np.dtype([('i', '<i4'), ('j', '<i4'), ('rank', '<i4'), ('slot', 'u1'), ('code', 'u1'),
                      ('box', '<u2'), ('cutoff_distance', '<f8'), ('max_gap', '<f8')])
BOX_DTYPE = np.dtype([('lim', '<f8', (12,)), ('name_id', '<i4'), ('lower_id', '<i4'), ('rev_lower_id', '<i4'),
                      ('subcat', '<i4'), ('name_len', '<i4'), ('flags', '<i4'), ('pad', '<i4', (2,))])
assert HIT_DTYPE.itemsize == 32 and BOX_DTYPE.itemsize == 128

_LIM_KEYS = ['xmin', 'xmax', 'ymin', 'ymax', 'zmin', 'zmax', 'normalmin', 'normalmax', 'anglemin', 'anglemax',
             'gapmax', 'radiusmax']


def reverse_edges(inter):
    """NA_pairwise_interactions.py:446-465."""
    if len(inter) <= 2 or inter == 'N/A':
        return inter
    if len(inter) == 3:
        return inter[0] + inter[2] + inter[1]
    if len(inter) == 4 and inter[0] in 'n!':
        return inter[0] + inter[1] + inter[3] + inter[2]
    if len(inter) == 4:
        return inter[0] + inter[2] + inter[1] + inter[3]
    if len(inter) == 5:
        return inter[0] + inter[1] + inter[3] + inter[2] + inter[4]
    return inter[0:(len(inter) - 2)] + inter[len(inter) - 1] + inter[len(inter) - 2]


class BoxTable(object):
    """Flattened focused_basepair_cutoffs: {combo: {+1/-1: {interaction: {subcategory: limits}}}}
    (the output of focus_basepair_cutoffs, NA:88-136) in iteration order."""

    def __init__(self, focused):
        rows = []
        self.names = []          # name_id -> interaction name
        self.box_name = []       # box -> interaction name
        self.box_subcat = []     # box -> subcategory (as given, may be any hashable)
        self.box_combo = []      # box -> 'A,U' parent combination it belongs to
        name_ids = {}
        lower_ids = {}
        index = np.zeros((_lib.N_PARENTS * _lib.N_PARENTS, 2, 2), dtype=np.int32)
        for a, pa in enumerate(defs.PARENTS):
            for b, pb in enumerate(defs.PARENTS):
                combo = pa + "," + pb
                if combo not in defs.BASEPAIR_COMBOS or combo not in focused:
                    continue
                for sg, key in enumerate((1, -1)):
                    inters = focused[combo].get(key, {})
                    start = len(rows)
                    for inter, subs in inters.items():
                        for sub, cut in subs.items():
                            row = np.zeros((), dtype=BOX_DTYPE)
                            for q, lk in enumerate(_LIM_KEYS):
                                if lk == 'radiusmax':
                                    row['lim'][q] = cut.get(lk, 0.0)
                                else:
                                    row['lim'][q] = cut[lk]
                            if inter not in name_ids:
                                name_ids[inter] = len(self.names)
                                self.names.append(inter)
                            row['name_id'] = name_ids[inter]
                            row['lower_id'] = lower_ids.setdefault(inter.lower(), len(lower_ids))
                            row['rev_lower_id'] = lower_ids.setdefault(reverse_edges(inter).lower(), len(lower_ids))
                            row['subcat'] = int(sub)
                            row['name_len'] = len(inter)
                            row['flags'] = (1 if inter.startswith("n") else 0) | (2 if inter == 'cSs' else 0) | \
                                           (4 if inter == 'csS' else 0) | (8 if 'radiusmax' in cut else 0)
                            rows.append(row)
                            self.box_name.append(inter)
                            self.box_subcat.append(sub)
                            self.box_combo.append(combo)
                    count = len(rows) - start
                    if count > 32:
                        raise _lib.Fr3dError("more than 32 cutoff boxes for %s sign %d" % (combo, key))
                    index[a * _lib.N_PARENTS + b, sg, 0] = start
                    index[a * _lib.N_PARENTS + b, sg, 1] = count
        if len(rows) > 65535:
            raise _lib.Fr3dError("too many cutoff boxes")
        self.boxes = np.array(rows, dtype=BOX_DTYPE) if rows else np.zeros(1, dtype=BOX_DTYPE)
        self.index = index
        self._dev = {}

    def device(self, torch, device):
        key = str(device)
        if key not in self._dev:
            bx = torch.from_numpy(self.boxes.view(np.uint8).reshape(-1).copy()).to(device)
            ix = torch.from_numpy(self.index.reshape(-1).copy()).to(device)
            self._dev[key] = (bx, ix)
        return self._dev[key]


class DeviceBatch(object):
    """The NtBatch arrays resident in HBM (torch tensors) + the fr3d_nts_t pointer block."""

    def __init__(self, torch, device, batch, pinned=None):
        self.n = batch.n
        self.compact_base = None
        t = {}
        for name in ("center", "rot", "base_xyz", "bb_xyz", "base_mask", "bb_mask", "meta"):
            if name in ("center", "rot") and not batch.frames_given:        # K1 writes them: nothing to upload
                shape = getattr(batch, name).shape
                t[name] = torch.empty(shape, dtype=torch.float64, device=device)
                continue
            src = pinned[name] if (pinned is not None and name in pinned) else torch.from_numpy(_as_signed(getattr(batch, name)))
            if name == "base_xyz" and src.shape[1] < _lib.BASE_SLOTS:       # compact pinned form, see pin_batch:
                # K1 reads the compact plane and writes the full 16-slot records (fr3d_frames_fit_from)
                self.compact_base = [src.to(device, non_blocking=True), int(src.shape[1]), None]
                t[name] = torch.empty((batch.n, _lib.BASE_SLOTS, 3), dtype=src.dtype, device=device)
            else:
                t[name] = src.to(device, non_blocking=True)
        if self.compact_base is not None:        # the uploaded mask plane stays as it is; K1 writes the record's mask
            self.compact_base[2] = t["base_mask"]
            t["base_mask"] = torch.empty_like(t["base_mask"])
        self.t = t
        self.nts = _lib.NtsStruct(n_nt=batch.n, reserved=0, center=t["center"].data_ptr(), rot=t["rot"].data_ptr(),
                                  base_xyz=t["base_xyz"].data_ptr(), bb_xyz=t["bb_xyz"].data_ptr(),
                                  base_mask=t["base_mask"].data_ptr(), bb_mask=t["bb_mask"].data_ptr(),
                                  meta=t["meta"].data_ptr())


def _as_signed(a):
    """torch has no uint32 arithmetic we need; move masks as int32 bit patterns."""
    if a.dtype == np.uint32:
        return a.view(np.int32)
    return a


HEAVY_SLOTS = 11        # the most heavy base atoms any parent has (G); slots from here on only ever hold hydrogens


def pinned_planes(batch):
    """The pinned torch tensors behind the record planes of a batch whose arrays already live in pinned memory
    (cif.native.cif_to_batch(..., pinned=True)), in the form Engine.annotate_batch(pinned=...) takes; None otherwise."""
    out = {}
    for name in ("base_xyz", "bb_xyz", "base_mask", "bb_mask", "meta"):
        t = getattr(batch, "_pin_" + name, None)
        if t is None:
            return None
        out[name] = t
    return out


def pin_batch(torch, batch):
    """Copy the batch into pinned host memory once (so timed H2D copies are DMA transfers).

    Hydrogens are normally not observed and are placed on the device by K1, so when no nt of the batch has an
    atom in base slots 11..15 the pinned base plane keeps only slots 0..10 (264 instead of 384 bytes per nt go
    over PCIe); the device record keeps all 16 slots."""
    out = {}
    for name in ("center", "rot", "base_xyz", "bb_xyz", "base_mask", "bb_mask", "meta"):
        a = _as_signed(getattr(batch, name))
        if name == "base_xyz" and batch.n and not (batch.base_mask & np.uint32(0xF800)).any():
            a = np.ascontiguousarray(a[:, :HEAVY_SLOTS])
        out[name] = torch.from_numpy(a).pin_memory()
    return out


class Engine(object):
    """One per process / GPU."""

    _instances = {}

    @classmethod
    def get(cls, device=None):
        import torch
        if not torch.cuda.is_available():
            raise _lib.Fr3dError("fr3d_python_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        key = dev.index if dev.index is not None else torch.cuda.current_device()
        if key not in cls._instances:
            cls._instances[key] = cls(torch, torch.device("cuda", key))
        return cls._instances[key]

    def __init__(self, torch, device):
        self.torch = torch
        self.device = device
        self.lib = _lib.load()
        _lib.check(self.lib.fr3d_init(device.index), "fr3d_init")
        self._tables = defs.static_tables()
        with torch.cuda.device(device):
            _lib.check(self.lib.fr3d_upload_tables(C.byref(self._tables)), "fr3d_upload_tables")
        self.launches = 0          # kernels launched by this engine (bench's gpu_launches)

    def stream_ptr(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    # ---------------------------------------------------------------- K1
    def _frames_fit(self, dbatch, status_ptr, place_h, st):
        compact = getattr(dbatch, "compact_base", None)
        packed = getattr(dbatch, "packed", None)
        if packed is not None:
            # the packed upload: backbone slots 0..8 + previous-O3' links -> bb_xyz, ragged observed base atoms -> K1
            _lib.check(self.lib.fr3d_backbone_expand(packed["bb9"].data_ptr(), packed["prev_rel"].data_ptr(), dbatch.n,
                                                     packed["extra_nt"].data_ptr(), packed["extra_xyz"].data_ptr(),
                                                     dbatch.packed_n_extra, dbatch.dev["bb_xyz"].data_ptr(), st),
                       "fr3d_backbone_expand")
            _lib.check(self.lib.fr3d_frames_fit_ragged(C.byref(dbatch.nts), packed["ragged"].data_ptr(),
                                                       packed["group_off"].data_ptr(), compact[2].data_ptr(), status_ptr,
                                                       1 if place_h else 0, st), "fr3d_frames_fit_ragged")
            self.launches += 1
        elif compact is not None:
            _lib.check(self.lib.fr3d_frames_fit_from(C.byref(dbatch.nts), compact[0].data_ptr(), compact[1],
                                                     compact[2].data_ptr(), status_ptr, 1 if place_h else 0, st),
                       "fr3d_frames_fit_from")
        else:
            _lib.check(self.lib.fr3d_frames_fit(C.byref(dbatch.nts), status_ptr, 1 if place_h else 0, st),
                       "fr3d_frames_fit")
        self.launches += 1

    def fit_frames(self, dbatch, place_h=True):
        torch = self.torch
        status = torch.empty(dbatch.n, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            self._frames_fit(dbatch, status.data_ptr(), place_h, self.stream_ptr())
        return status

    # ---------------------------------------------------------------- plans
    def make_plan(self, n_nt, pair_cap=None, hit_cap=None, tail=None, row_cap=None, row_out=None):
        """Pre-allocate every device buffer one pass over an n_nt batch needs, so that the pass
        itself is launches only (no allocation, no host sync).  tail = (n_struct, n_ends): also the buffers of the
        device tail (fr3d_annotation_tail)."""
        torch = self.torch
        if pair_cap is None:
            pair_cap = 12 * n_nt + 1024        # ~6.3 candidate pairs per nt on RNA
        if hit_cap is None:
            hit_cap = pair_cap + 1024
        p = Plan()
        p.n_nt, p.pair_cap, p.hit_cap = n_nt, int(pair_cap), int(hit_cap)
        p.ws_bytes = int(self.lib.fr3d_pair_search_workspace(n_nt))
        dev = self.device
        p.ws = torch.empty(p.ws_bytes, dtype=torch.uint8, device=dev)
        p.pair_i = torch.empty(p.pair_cap, dtype=torch.int32, device=dev)
        p.pair_j = torch.empty(p.pair_cap, dtype=torch.int32, device=dev)
        p.pair_rank = torch.empty(p.pair_cap, dtype=torch.int32, device=dev)
        p.counters = torch.zeros(8, dtype=torch.int64, device=dev)
        p.cws_bytes = int(self.lib.fr3d_classify_workspace(p.pair_cap, n_nt))
        p.cws = torch.empty(p.cws_bytes, dtype=torch.uint8, device=dev)
        p.hits = torch.empty(p.hit_cap * HIT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        p.status = torch.zeros(max(n_nt, 1), dtype=torch.int32, device=dev)
        p.tail = None
        if tail is not None:
            self.add_tail_buffers(p, tail[0], tail[1], row_cap, out=row_out)
        return p

    def add_tail_buffers(self, p, n_struct, n_ends, row_cap=None, out=None):
        """out = (rows uint8 [16 row_cap], row_start int32 [n_struct], row_count int32 [n_struct], tcounters int64
        [32]): caller-owned output tensors (e.g. views of a buffer a collective sends) instead of new ones."""
        torch, dev = self.torch, self.device
        if row_cap is None:
            row_cap = 3 * p.hit_cap + 2 * p.n_nt + 1024      # ~2.2 rows per hit + one covalent row per nt
        p.tail = (int(n_struct), int(n_ends))
        p.row_cap = int(row_cap)
        p.tws_bytes = int(self.lib.fr3d_tail_workspace(p.hit_cap, p.n_nt, p.tail[0], p.tail[1]))
        p.tws = torch.empty(p.tws_bytes, dtype=torch.uint8, device=dev)
        p.external_rows = out is not None
        if out is not None:
            p.rows, p.row_start, p.row_count, p.tcounters = out
            assert p.rows.numel() >= 16 * p.row_cap and p.row_start.numel() >= p.tail[0] and p.tcounters.numel() >= 32
            return
        p.rows = torch.empty(p.row_cap * 16, dtype=torch.uint8, device=dev)
        p.row_start = torch.empty(max(p.tail[0], 1), dtype=torch.int32, device=dev)
        p.row_count = torch.empty(max(p.tail[0], 1), dtype=torch.int32, device=dev)
        p.tcounters = torch.zeros(32, dtype=torch.int64, device=dev)      # [0..3] see fr3d_annotation_tail; the rest: profiling builds

    def run_plan(self, dbatch, plan, boxtable, category_mask, fit_frames=True, events=None, index_offset=0, tail=None):
        """K1 -> K2/K3 -> K4 [-> device tail] on the current stream, asynchronously.  `events` (dict) receives
        torch CUDA events around the classify launch for per-kernel timing.  tail = (DeviceIdent, LabelTable): also
        run fr3d_annotation_tail into the plan's row buffers."""
        torch = self.torch
        bx, ix = boxtable.device(torch, self.device)
        st = self.stream_ptr()
        with torch.cuda.device(self.device):
            if fit_frames:
                self._frames_fit(dbatch, plan.status.data_ptr(), True, st)
            if events is not None:
                events["search0"].record()
            _lib.check(self.lib.fr3d_pair_search(C.byref(dbatch.nts), 12.0, plan.ws.data_ptr(), plan.ws_bytes,
                                                 plan.pair_i.data_ptr(), plan.pair_j.data_ptr(),
                                                 plan.pair_rank.data_ptr(), plan.pair_cap, plan.counters.data_ptr(), st),
                       "fr3d_pair_search")
            self.launches += 5
            if events is not None:
                events["classify0"].record()
            _lib.check(self.lib.fr3d_classify_pairs(C.byref(dbatch.nts), bx.data_ptr(), ix.data_ptr(),
                                                    plan.pair_i.data_ptr(), plan.pair_j.data_ptr(),
                                                    plan.pair_rank.data_ptr(), plan.pair_cap, int(category_mask),
                                                    int(index_offset), plan.cws.data_ptr(), plan.cws_bytes,
                                                    plan.hits.data_ptr(), plan.hit_cap, plan.counters.data_ptr(), st),
                       "fr3d_classify_pairs")
            # screen records + screen + one thread-per-pair kernel per work list (lists of families that were not
            # requested are skipped)
            cm = int(category_mask)
            self.launches += 3 + (1 if cm & (_lib.CAT_SO | _lib.CAT_SUGAR_RIBOSE) else 0) + \
                (1 if cm & _lib.CAT_BACKBONE else 0) + (2 if cm & _lib.CAT_STACKING else 0)
            if events is not None:
                events["classify1"].record()
            if tail is not None:
                self.run_tail(dbatch, plan, tail, index_offset, st)
                if events is not None and "tail1" in events:
                    events["tail1"].record()

    def run_tail(self, dbatch, plan, tail, index_offset=0, st=None):
        """fr3d_annotation_tail on the hit records of `plan` (5 launches)."""
        ident, labels = tail
        tables, _keep = labels.device(self.torch, self.device)
        if st is None:
            st = self.stream_ptr()
        _lib.check(self.lib.fr3d_annotation_tail(C.byref(dbatch.nts), C.byref(ident.struct), C.byref(tables),
                                                 plan.hits.data_ptr(), plan.hit_cap, plan.counters.data_ptr(),
                                                 int(index_offset), plan.tws.data_ptr(), plan.tws_bytes,
                                                 plan.rows.data_ptr(), plan.row_cap, plan.row_start.data_ptr(),
                                                 plan.row_count.data_ptr(), plan.tcounters.data_ptr(), st),
                   "fr3d_annotation_tail")
        self.launches += 5 if ident.n_struct > 0 else 1

    def read_counts(self, plan):
        """(n_pairs, n_hits) after a sync; raises on range errors; None if a capacity was exceeded."""
        cnt = plan.counters.cpu()
        if int(cnt[2]) != 0:
            raise _lib.Fr3dError("fr3d_pair_search: base centre outside the spatial-hash range (FR3D_E_RANGE)")
        return int(cnt[0]), int(cnt[1])

    # ---------------------------------------------------------------- whole path
    def annotate_batch(self, batch, boxtable, category_mask, pinned=None, return_frames=False, plan=None,
                       labels=None, want_hits=True):
        """H2D -> [K1] -> K2/K3 -> K4 [-> device tail] -> D2H.  Returns (hits structured array, n_pairs, extras).
        With `labels` (a tail.LabelTable) the device tail runs too and extras["rows"], ["row_start"], ["row_count"]
        hold the final rows (hits are then only copied back if want_hits).  Capacity overflow (FR3D_E_CAPACITY)
        re-runs with the exact counts; nothing is truncated."""
        torch = self.torch
        with torch.cuda.device(self.device):
            db = DeviceBatch(torch, self.device, batch, pinned)
            tail = None
            if labels is not None:
                from . import tail as tailmod
                ident = DeviceIdent(torch, self.device, tailmod.tail_ident(batch))
                tail = (ident, labels)
            if plan is None or plan.n_nt != batch.n:
                plan = self.make_plan(batch.n)
            if tail is not None and plan.tail != (ident.n_struct, ident.n_ends):
                self.add_tail_buffers(plan, ident.n_struct, ident.n_ends)
            fit = not batch.frames_given
            status = None
            while True:
                self.run_plan(db, plan, boxtable, category_mask, fit_frames=fit, tail=tail)
                if fit:
                    status = plan.status        # K1 ran in this pass: its status survives capacity re-runs
                n_pairs, n_hits = self.read_counts(plan)
                if n_pairs <= plan.pair_cap and n_hits <= plan.hit_cap:
                    break
                fit = False          # frames are already in place in the device batch
                plan = self.make_plan(batch.n, max(n_pairs, plan.pair_cap) + 1024,
                                      max(n_hits, 2 * max(n_pairs, 1)) + 1024,
                                      tail=(ident.n_struct, ident.n_ends) if tail is not None else None)
            extras = {"plan": plan}
            if tail is not None:
                while True:
                    n_rows = self.read_tail_counts(plan)
                    if n_rows <= plan.row_cap:
                        break
                    self.add_tail_buffers(plan, ident.n_struct, ident.n_ends, row_cap=n_rows + 1024)
                    self.run_tail(db, plan, tail)
                from .tail import ROW_DTYPE
                extras["rows"] = plan.rows[: n_rows * 16].cpu().numpy().view(ROW_DTYPE)
                extras["row_start"] = plan.row_start[:ident.n_struct].cpu().numpy()
                extras["row_count"] = plan.row_count[:ident.n_struct].cpu().numpy()
            if want_hits or tail is None:
                hits = plan.hits[: n_hits * HIT_DTYPE.itemsize].cpu().numpy().view(HIT_DTYPE)
            else:
                hits = None
            extras["n_hits"] = n_hits
            if return_frames:
                extras["center"] = db.t["center"].cpu().numpy()
                extras["rot"] = db.t["rot"].cpu().numpy()
                extras["base_xyz"] = db.t["base_xyz"].cpu().numpy()
                extras["base_mask"] = db.t["base_mask"].cpu().numpy().view(np.uint32)
                extras["status"] = status[:batch.n].cpu().numpy() if status is not None else None
                extras["pairs"] = (plan.pair_i[:n_pairs].cpu().numpy(), plan.pair_j[:n_pairs].cpu().numpy(),
                                   plan.pair_rank[:n_pairs].cpu().numpy())
        return hits, n_pairs, extras

    def read_tail_counts(self, plan):
        """Total rows after a sync; raises when a structure is outside the ranges the device tail supports."""
        tc = plan.tcounters.cpu()
        if int(tc[2]) != 0:
            raise _lib.Fr3dError("fr3d_annotation_tail: a structure is outside the supported ranges (more than 131 072 "
                                 "nts, an index outside [0, 2^23), or a label without its table entry): FR3D_E_RANGE")
        return int(tc[0])


class DeviceIdent(object):
    """The fr3d_tail_ident_t arrays of a batch (or of structures s0..s1-1 of it) in HBM."""

    def __init__(self, torch, device, ident, s0=0, s1=None, pinned=False):
        if s1 is None:
            s1 = ident.n_struct
        arrays = ident.chunk(s0, s1)
        self.n_struct, self.n_chains, self.n_ends = s1 - s0, arrays[5], arrays[6]
        self.host = arrays[:5]
        # Large structures (>= 1 024 nts) get the tail's 1 024-thread blocks, one per SM: the shortest time for ONE structure.
        # A batch of MANY of them is served better by the ordinary 256-thread blocks, four per SM (the heavy bench batch,
        # 1000 x 2 844 nt: tail 3.6 ms with the large blocks, 2.8 ms without), so they are only announced when they are few.
        sizes = np.diff(arrays[0]) if s1 > s0 else np.zeros(0, dtype=np.int64)
        n_big = int((sizes >= 1024).sum())
        limit = int(os.environ.get("FR3D_TAIL_BIG_LIMIT", "160"))      # ~ one block per SM
        self.max_struct_nt = int(sizes.max()) if (len(sizes) and 0 < n_big <= limit) else 0
        offs, total = [], 0
        for a in self.host:
            offs.append(total)
            total += (a.nbytes + 255) // 256 * 256
        self.offs = offs
        blob = np.zeros(max(total, 1), dtype=np.uint8)
        for o, a in zip(offs, self.host):
            blob[o:o + a.nbytes] = a.view(np.uint8).reshape(-1)
        self.host_blob = torch.from_numpy(blob)
        if pinned:
            self.host_blob = self.host_blob.pin_memory()
        self.dev_blob = None
        if device is not None:
            self.bind(self.host_blob.to(device, non_blocking=pinned))

    def bind(self, dev_blob):
        """Point the struct at a device copy of host_blob (the pipeline uploads it with the nt records)."""
        self.dev_blob = dev_blob
        base = dev_blob.data_ptr()
        o = self.offs
        self.struct = _lib.TailIdentStruct(n_struct=self.n_struct, n_chains=self.n_chains, n_ends=self.n_ends,
                                           max_struct_nt=self.max_struct_nt, struct_off=base + o[0], struct_chain_off=base + o[1],
                                           chain_ends_off=base + o[2], nt_chain=base + o[3], nt_cov=base + o[4])


class Plan(object):
    pass


def chunk_bounds(S, n_chunks, weights=None):
    """Structure boundaries of the pipeline chunks: equal parts, or parts proportional to `weights` (one per chunk) with at
    least one structure each."""
    n_chunks = max(1, min(int(n_chunks), int(S))) if S > 0 else 1
    if weights is not None and len(weights) == n_chunks and n_chunks > 1:
        w = np.maximum(np.asarray(weights, dtype=float), 0.0)
        cum = np.concatenate([[0.0], np.cumsum(w)]) / max(float(w.sum()), 1e-300)
        bounds = [int(round(c * S)) for c in cum]
        for k in range(1, n_chunks + 1):                      # every chunk keeps at least one structure
            bounds[k] = min(max(bounds[k], bounds[k - 1] + 1), S - (n_chunks - k))
        return bounds
    return [int(round(k * S / n_chunks)) for k in range(n_chunks + 1)]


def _pack_upload(base, base_mask, bb, bb_mask):
    """The packed upload of one chunk (host numpy): base [n][11][3] with its observed-mask plane, bb [n][10][3] with its mask
    -> ragged [A][3] (observed base atoms only, nt by nt in slot order), group_off [ceil(n / 32) + 1] (first atom of every
    32 nts), bb9 [n][9][3], prev_rel int8 [n] (1: slot 9 is the previous ROW's O3'), and the explicit (nt, xyz) list of the
    previous-O3' entries that are not."""
    n = base.shape[0]
    bits = ((base_mask.astype(np.uint32)[:, None] >> np.arange(base.shape[1], dtype=np.uint32)) & 1).astype(bool)
    ragged = np.ascontiguousarray(base[bits])
    off = np.concatenate([[0], np.cumsum(bits.sum(1))]).astype(np.int64)
    group_off = off[np.minimum(np.arange(0, n + 32, 32), n)].astype(np.int32)
    if group_off.shape[0] != (n + 31) // 32 + 1:
        group_off = group_off[:(n + 31) // 32 + 1]
    has9 = ((bb_mask >> 9) & 1).astype(bool)
    prev_ok = np.zeros(n, dtype=bool)
    if n > 1:
        prev_ok[1:] = has9[1:] & ((bb_mask[:-1] >> 5) & 1).astype(bool) & np.all(bb[1:, 9] == bb[:-1, 5], axis=1)
    extra = np.nonzero(has9 & ~prev_ok)[0].astype(np.int32)
    return {"ragged": ragged if ragged.size else np.zeros((1, 3)), "group_off": group_off,
            "bb9": np.ascontiguousarray(bb[:, :9]), "prev_rel": prev_ok.astype(np.int8),
            "extra_nt": extra, "extra_xyz": np.ascontiguousarray(bb[extra, 9])}


_FIELDS = ("center", "rot", "base_xyz", "bb_xyz", "base_mask", "bb_mask", "meta")


class Pipeline(object):
    """Chunked, multi-stream form of Engine.annotate_batch for large batches: structures are
    independent, so the batch is cut into contiguous chunks of whole structures and chunk k+1's
    H2D copy overlaps chunk k's kernels and chunk k-1's D2H copy.  All device / pinned buffers are
    allocated once and re-used across calls.  Frames are always fitted on the device (K1), so the
    centre / rotation planes are not uploaded.  The input is fixed at construction: each chunk's planes are laid out
    back to back in one pinned blob here, and run() uploads those blobs (build a new Pipeline for new coordinates).

    With `labels` (a tail.LabelTable) every chunk also runs the device tail and run() returns the final rows;
    the hit records are then only copied back when want_hits is set."""

    def __init__(self, engine, batch, pinned, boxtable, category_mask, n_chunks=4, pair_cap=None, hit_cap=None,
                 n_streams=3, labels=None, want_hits=None, row_cap=None, chunk_caps=None, row_out=None, chunk_weights=None,
                 packed_upload=True):
        """packed_upload (with a compact base plane): the upload carries only the OBSERVED base atoms (ragged, ~9.3 of 11
        slots per nt) and backbone slots 0..8 - slot 9, the previous residue's O3', is rebuilt on the device from the
        previous row - 490 instead of 552 bytes per nt over PCIe; K1 (fr3d_frames_fit_ragged) and fr3d_backbone_expand write
        the ordinary records.  chunk_weights: relative sizes of the chunks (default: equal).  The first chunk's upload and the last chunk's
        kernels + download are the parts of a pass nothing overlaps with, so a tapered split (small first and last chunks)
        shortens the pass.  chunk_caps: per chunk (pair_cap, hit_cap, row_cap) instead of the proportional split of pair_cap / hit_cap /
        row_cap; row_out(k, n_struct, row_cap) -> (rows, row_start, row_count, tcounters): caller-owned output tensors
        of chunk k (see Engine.add_tail_buffers), for enqueue() + a collective."""
        torch = engine.torch
        self.eng, self.batch, self.pinned, self.bt, self.mask = engine, batch, pinned, boxtable, category_mask
        self.labels = labels
        self.want_hits = (labels is None) if want_hits is None else bool(want_hits)
        if batch.frames_given:
            raise _lib.Fr3dError("Pipeline uploads observed atoms only; frames are fitted on the device")
        ident = None
        if labels is not None:
            from . import tail as tailmod
            ident = tailmod.tail_ident(batch)
        S = batch.n_structures
        n_chunks = max(1, min(n_chunks, S))
        bounds = chunk_bounds(S, n_chunks, chunk_weights)
        self.chunks = []
        dev = engine.device
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(min(n_streams, n_chunks))]
        self.h2d_bytes = 0
        for k in range(n_chunks):
            lo, hi = int(batch.struct_off[bounds[k]]), int(batch.struct_off[bounds[k + 1]])
            n = hi - lo
            ch = Plan()
            ch.lo, ch.hi, ch.n = lo, hi, n
            ch.s0, ch.s1 = bounds[k], bounds[k + 1]
            ch.stream = self.streams[k % len(self.streams)]
            # One host -> device copy per chunk: the chunk's slices of the uploaded planes (and the identity arrays
            # of the tail) sit back to back in ONE pinned blob (built here, once) and land in ONE device blob whose
            # sub-ranges ARE the device planes (PCIe moves 40 MB pieces at ~55 GB/s but 2-10 MB pieces at ~36 GB/s,
            # profiles/pcie_probe.json).  A compact base plane (slots 0..10) is read by K1 directly, which writes
            # the full 16-slot records.
            compact = pinned["base_xyz"].shape[1] < _lib.BASE_SLOTS
            uploaded = [name for name in _FIELDS if name not in ("center", "rot")]
            packed = None
            if compact and packed_upload and n > 0:
                packed = _pack_upload(pinned["base_xyz"][lo:hi].numpy(), pinned["base_mask"][lo:hi].numpy(),
                                      pinned["bb_xyz"][lo:hi].numpy(), pinned["bb_mask"][lo:hi].numpy())
                uploaded = [name for name in uploaded if name not in ("base_xyz", "bb_xyz")]
            offs, total = {}, 0
            for name in uploaded:
                src = pinned[name][lo:hi]
                offs[name] = (total, src.numel() * src.element_size(), src.dtype, tuple(src.shape))
                total += (offs[name][1] + 255) // 256 * 256
            poffs = {}
            if packed is not None:
                for name, arr in packed.items():
                    poffs[name] = (total, arr.nbytes)
                    total += (arr.nbytes + 255) // 256 * 256
            ch.ident = None
            if ident is not None:
                ch.ident = DeviceIdent(torch, None, ident, ch.s0, ch.s1)
                ident_off = total
                total += (ch.ident.host_blob.numel() + 255) // 256 * 256
            ch.host_blob = torch.empty(max(total, 1), dtype=torch.uint8).pin_memory()
            ch.dev_blob = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
            self.h2d_bytes += total
            ch.dev = {}
            ch.compact_base = None
            for name in uploaded:
                o, nb, dt, shape = offs[name]
                ch.host_blob[o:o + nb].view(dt).view(shape).copy_(pinned[name][lo:hi])
                view = ch.dev_blob[o:o + nb].view(dt).view(shape)
                if name == "base_xyz" and compact:
                    ch.compact_base = [view, int(shape[1]), None]
                    ch.dev[name] = torch.empty((n, _lib.BASE_SLOTS, 3), dtype=dt, device=dev)
                elif name == "base_mask" and compact and packed is None:
                    ch.compact_base[2] = view          # the uploaded (observed) masks; K1 writes the record's own
                    ch.dev[name] = torch.empty((n,), dtype=dt, device=dev)
                else:
                    ch.dev[name] = view
            ch.packed = None
            if packed is not None:
                views = {}
                for name, arr in packed.items():
                    o, nb = poffs[name]
                    if nb:
                        ch.host_blob[o:o + nb].copy_(torch.from_numpy(arr.view(np.uint8).reshape(-1)))
                    views[name] = ch.dev_blob[o:o + max(nb, 1)]
                ch.packed = views
                ch.packed_n_extra = int(packed["extra_nt"].shape[0])
                ch.compact_base = [views["ragged"], 11, ch.dev_blob[offs["base_mask"][0]:offs["base_mask"][0] + offs["base_mask"][1]]
                                   .view(pinned["base_mask"].dtype)]
                ch.dev["base_xyz"] = torch.empty((n, _lib.BASE_SLOTS, 3), dtype=pinned["base_xyz"].dtype, device=dev)
                ch.dev["base_mask"] = torch.empty((n,), dtype=pinned["base_mask"].dtype, device=dev)
                ch.dev["bb_xyz"] = torch.empty((n, _lib.BB_SLOTS, 3), dtype=pinned["bb_xyz"].dtype, device=dev)
            if ch.ident is not None:
                nb = ch.ident.host_blob.numel()
                ch.host_blob[ident_off:ident_off + nb].copy_(ch.ident.host_blob)
                ch.ident.bind(ch.dev_blob[ident_off:ident_off + nb])
            for name in ("center", "rot"):
                ch.dev[name] = torch.zeros((n,) + tuple(getattr(batch, name).shape[1:]), dtype=pinned[name].dtype,
                                           device=dev)
            t = ch.dev
            ch.nts = _lib.NtsStruct(n_nt=n, reserved=0, center=t["center"].data_ptr(), rot=t["rot"].data_ptr(),
                                    base_xyz=t["base_xyz"].data_ptr(), bb_xyz=t["bb_xyz"].data_ptr(),
                                    base_mask=t["base_mask"].data_ptr(), bb_mask=t["bb_mask"].data_ptr(),
                                    meta=t["meta"].data_ptr())
            frac = n / max(batch.n, 1)
            pc = int((pair_cap or (12 * batch.n + 1024)) * frac * 1.15) + 4096
            hc = int((hit_cap or (12 * batch.n + 1024)) * frac * 1.15) + 4096
            rc = (int(row_cap * frac * 1.15) + 4096) if row_cap else None
            if chunk_caps is not None:
                pc, hc, rc = chunk_caps[k]
            ro = row_out(k, ch.ident.n_struct, rc) if (row_out is not None and ch.ident is not None) else None
            ch.plan = engine.make_plan(n, pc, hc, tail=(ch.ident.n_struct, ch.ident.n_ends) if ch.ident else None,
                                       row_cap=rc, row_out=ro)
            ch.counters_host = torch.zeros(12, dtype=torch.int64).pin_memory()
            ch.ev = torch.cuda.Event()
            self.chunks.append(ch)
        self._alloc_results()

    def _alloc_results(self):
        torch = self.eng.torch
        self.hits_host = None
        if self.want_hits:
            total_hc = sum(ch.plan.hit_cap for ch in self.chunks)
            self.hits_host = torch.empty(total_hc * HIT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
        self.rows_host = None
        if self.labels is not None:
            total_rc = sum(ch.plan.row_cap for ch in self.chunks)
            S = self.batch.n_structures
            self.rows_host = torch.empty(total_rc * 16, dtype=torch.uint8).pin_memory()
            self.row_start_host = torch.empty(max(S, 1), dtype=torch.int32).pin_memory()
            self.row_count_host = torch.empty(max(S, 1), dtype=torch.int32).pin_memory()

    def _launch(self, ch):
        eng = self.eng
        tail = (ch.ident, self.labels) if ch.ident is not None else None
        eng.run_plan(ch, ch.plan, self.bt, self.mask, fit_frames=True, index_offset=ch.lo, tail=tail)
        ch.counters_host[:8].copy_(ch.plan.counters, non_blocking=True)
        if tail is not None:
            ch.counters_host[8:].copy_(ch.plan.tcounters[:4], non_blocking=True)
        ch.ev.record(ch.stream)

    def build_graph(self):
        """Capture one whole pass - for every chunk: H2D, K1..K7, counters, and the D2H of its row block at full
        capacity into a fixed region of the pinned result buffer, chunks forked onto their streams and joined -
        into ONE CUDA graph, so that a pass costs one graph launch instead of ~30 launches and copies per chunk from
        Python (the pipeline is launch-bound on the host once the kernels are fast).  Needs the device tail."""
        eng, torch = self.eng, self.eng.torch
        if self.labels is None:
            raise _lib.Fr3dError("Pipeline.build_graph needs the device tail (labels)")
        self.run()                                   # warm-up; also grows any capacity that was too small
        self._alloc_results()
        side = torch.cuda.Stream(device=eng.device)
        g = torch.cuda.CUDAGraph()
        launches0 = eng.launches
        torch.cuda.synchronize(eng.device)
        with torch.cuda.graph(g, stream=side):
            main = torch.cuda.current_stream(eng.device)
            start = torch.cuda.Event()
            start.record(main)
            roff = 0
            for ch in self.chunks:
                ch.stream.wait_event(start)
                with torch.cuda.stream(ch.stream):
                    ch.dev_blob.copy_(ch.host_blob, non_blocking=True)
                    self._launch(ch)
                    nb = ch.plan.row_cap * 16
                    ns = ch.s1 - ch.s0
                    self.rows_host[roff:roff + nb].copy_(ch.plan.rows[:nb], non_blocking=True)
                    self.row_start_host[ch.s0:ch.s1].copy_(ch.plan.row_start[:ns], non_blocking=True)
                    self.row_count_host[ch.s0:ch.s1].copy_(ch.plan.row_count[:ns], non_blocking=True)
                    ch.graph_row_off = roff // 16
                    done = torch.cuda.Event()
                    done.record(ch.stream)
                main.wait_event(done)
                roff += nb
        self.graph = g
        self.graph_launches = eng.launches - launches0
        self.graph_d2h_bytes = roff + 8 * self.batch.n_structures + 96 * len(self.chunks)
        return g

    def run_graph(self):
        """One pass by replaying the captured graph (build_graph() on first use).  Same result object as run(); the
        row array has one capacity-sized region per chunk (row_start points into it)."""
        eng, torch = self.eng, self.eng.torch
        if getattr(self, "graph", None) is None:
            self.build_graph()
        self.graph.replay()
        torch.cuda.current_stream(eng.device).synchronize()
        eng.launches += self.graph_launches
        n_pairs = n_hits = 0
        for ch in self.chunks:
            cnt = ch.counters_host
            if int(cnt[2]) != 0:
                raise _lib.Fr3dError("fr3d_pair_search: base centre outside the spatial-hash range (FR3D_E_RANGE)")
            npairs, nhits, nrows = int(cnt[0]), int(cnt[1]), int(cnt[8])
            if npairs > ch.plan.pair_cap or nhits > ch.plan.hit_cap or nrows > ch.plan.row_cap:
                self.graph = None                    # capacities of the captured buffers: grow them (run) and re-capture
                self.build_graph()
                return self.run_graph()
            if int(cnt[10]) != 0:
                raise _lib.Fr3dError("fr3d_annotation_tail: a structure is outside the supported ranges (FR3D_E_RANGE)")
            ch.n_hits, ch.n_rows = nhits, nrows
            n_pairs += npairs
            n_hits += nhits
        from .tail import ROW_DTYPE
        S = self.batch.n_structures
        res = PipelineResult()
        res.n_pairs, res.n_hits, res.hits = n_pairs, n_hits, None
        res.rows = self.rows_host.numpy().view(ROW_DTYPE)
        res.row_start = self.row_start_host.numpy()[:S].astype(np.int64)
        res.row_count = self.row_count_host.numpy()[:S]
        res.n_rows = int(sum(ch.n_rows for ch in self.chunks))
        for ch in self.chunks:
            res.row_start[ch.s0:ch.s1] += ch.graph_row_off
        return res

    def enqueue_graph(self, upload=True, extra=None):
        """enqueue() as ONE CUDA graph launch on the current stream (captured on first use, one graph per value of
        `upload`); extra(): more work on the current stream to capture behind the chunks (e.g. copying counters into
        a send block).  Falls back to plain enqueue() + extra() if the capture fails."""
        eng, torch = self.eng, self.eng.torch
        graphs = self.__dict__.setdefault("_enqueue_graphs", {})
        key = bool(upload)
        if key not in graphs:
            self.enqueue(upload=upload)                     # warm-up outside the capture
            if extra is not None:
                extra()
            torch.cuda.synchronize(eng.device)
            launches0 = eng.launches
            try:
                g = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream(device=eng.device)
                with torch.cuda.graph(g, stream=side):
                    self.enqueue(upload=upload)
                    if extra is not None:
                        extra()
                graphs[key] = (g, eng.launches - launches0)
            except Exception:
                graphs[key] = None
        ent = graphs[key]
        if ent is None:
            self.enqueue(upload=upload)
            if extra is not None:
                extra()
            return
        ent[0].replay()
        eng.launches += ent[1]

    def enqueue(self, upload=True):
        """Issue one pass - per chunk: H2D (unless upload is False: the records of the previous pass are still in
        HBM), K1..K4, tail - on the chunk streams and make the CURRENT stream wait for
        all of it; no host synchronisation, nothing copied back.  For callers that own the output buffers (row_out)
        and follow up with a collective on the current stream; capacities are theirs to check (plan.counters,
        plan.tcounters travel with the results)."""
        eng, torch = self.eng, self.eng.torch
        main = torch.cuda.current_stream(eng.device)
        start = torch.cuda.Event()
        start.record(main)
        for ch in self.chunks:
            ch.stream.wait_event(start)
            with torch.cuda.stream(ch.stream):
                if upload or (ch.compact_base is None):        # in-place records are completed by K1: always re-upload
                    ch.dev_blob.copy_(ch.host_blob, non_blocking=True)
                tail = (ch.ident, self.labels) if ch.ident is not None else None
                eng.run_plan(ch, ch.plan, self.bt, self.mask, fit_frames=True, index_offset=ch.lo, tail=tail)
                ch.ev.record(ch.stream)
            main.wait_event(ch.ev)

    def run(self, to_host=True):
        """One pass.  Returns a PipelineResult: n_pairs, n_hits, and (with labels) rows / row_start / row_count -
        structured-array VIEWS of the pipeline's pinned result buffers with GLOBAL nt indices, valid until the
        next run(); .hits when the pipeline copies hit records back.  A chunk whose capacities turn out too small
        is re-run with the exact counts (nothing is truncated).
        to_host=False (needs labels): nothing is copied back; the result holds DEVICE tensors instead - rows_dev
        int32 [n_rows, 4] (the chunks' blocks concatenated), row_start_dev / row_count_dev int32 [n_structures] -
        for a collective to pick up (parallel.ShardedAnnotator)."""
        eng, torch = self.eng, self.eng.torch
        main = torch.cuda.current_stream(eng.device)
        start = torch.cuda.Event()
        start.record(main)
        for ch in self.chunks:
            ch.stream.wait_event(start)
            with torch.cuda.stream(ch.stream):
                ch.dev_blob.copy_(ch.host_blob, non_blocking=True)
                self._launch(ch)
        n_pairs = n_hits = 0
        hoff = roff = 0
        regrow = False
        for ch in self.chunks:
            while True:
                ch.ev.synchronize()
                cnt = ch.counters_host
                if int(cnt[2]) != 0:
                    raise _lib.Fr3dError("fr3d_pair_search: base centre outside the spatial-hash range (FR3D_E_RANGE)")
                npairs, nhits = int(cnt[0]), int(cnt[1])
                nrows = int(cnt[8]) if ch.ident is not None else 0
                if ch.ident is not None and int(cnt[10]) != 0 and npairs <= ch.plan.pair_cap and nhits <= ch.plan.hit_cap:
                    raise _lib.Fr3dError("fr3d_annotation_tail: a structure is outside the supported ranges (FR3D_E_RANGE)")
                over = npairs > ch.plan.pair_cap or nhits > ch.plan.hit_cap
                if not over and (ch.ident is None or nrows <= ch.plan.row_cap):
                    break
                # capacity exceeded: new buffers with the exact counts, this chunk again (its input is still on the device)
                if getattr(ch.plan, "external_rows", False):
                    raise _lib.Fr3dError("pipeline chunk capacity exceeded (pairs %d/%d, hits %d/%d, rows %d/%d) with "
                                         "caller-owned row buffers" % (npairs, ch.plan.pair_cap, nhits, ch.plan.hit_cap,
                                                                       nrows, ch.plan.row_cap))
                regrow = True
                tail = (ch.ident.n_struct, ch.ident.n_ends) if ch.ident is not None else None
                if over:
                    ch.plan = eng.make_plan(ch.n, max(npairs, ch.plan.pair_cap) + 1024,
                                            max(nhits, 2 * npairs if nhits > ch.plan.hit_cap else ch.plan.hit_cap) + 1024,
                                            tail=tail, row_cap=None)
                else:
                    eng.add_tail_buffers(ch.plan, tail[0], tail[1], row_cap=nrows + 1024)
                with torch.cuda.stream(ch.stream):
                    if ch.compact_base is None:
                        ch.dev_blob.copy_(ch.host_blob, non_blocking=True)     # K1 has completed the records in place
                    self._launch(ch)
            ch.n_hits, ch.n_rows = nhits, nrows
            n_pairs += npairs
            n_hits += nhits
        if regrow:
            self._alloc_results()
        if not to_host:
            if self.labels is None:
                raise _lib.Fr3dError("Pipeline.run(to_host=False) needs the device tail (labels)")
            res = PipelineResult()
            res.n_pairs, res.n_hits = n_pairs, n_hits
            parts, starts, counts, off = [], [], [], 0
            for ch in self.chunks:                # all chunk kernels have finished (events), so any stream may read
                ns = ch.s1 - ch.s0
                parts.append(ch.plan.rows[:ch.n_rows * 16].view(torch.int32).view(-1, 4))
                starts.append(ch.plan.row_start[:ns] + off)
                counts.append(ch.plan.row_count[:ns])
                off += ch.n_rows
            res.rows_dev = torch.cat(parts) if parts else torch.zeros((0, 4), dtype=torch.int32, device=eng.device)
            res.row_start_dev = torch.cat(starts) if starts else torch.zeros(0, dtype=torch.int32, device=eng.device)
            res.row_count_dev = torch.cat(counts) if counts else torch.zeros(0, dtype=torch.int32, device=eng.device)
            res.n_rows = off
            return res
        for ch in self.chunks:
            with torch.cuda.stream(ch.stream):
                if self.hits_host is not None:
                    nb = ch.n_hits * HIT_DTYPE.itemsize
                    self.hits_host[hoff:hoff + nb].copy_(ch.plan.hits[:nb], non_blocking=True)
                    hoff += nb
                if ch.ident is not None:
                    nb = ch.n_rows * 16
                    self.rows_host[roff:roff + nb].copy_(ch.plan.rows[:nb], non_blocking=True)
                    ns = ch.s1 - ch.s0
                    self.row_start_host[ch.s0:ch.s1].copy_(ch.plan.row_start[:ns], non_blocking=True)
                    self.row_count_host[ch.s0:ch.s1].copy_(ch.plan.row_count[:ns], non_blocking=True)
                    ch.row_off = roff // 16
                    roff += nb
        for st in self.streams:
            st.synchronize()
        res = PipelineResult()
        res.n_pairs, res.n_hits = n_pairs, n_hits
        res.hits = self.hits_host.numpy()[:hoff].view(HIT_DTYPE) if self.hits_host is not None else None
        res.rows = res.row_start = res.row_count = None
        if self.labels is not None:
            from .tail import ROW_DTYPE
            S = self.batch.n_structures
            res.rows = self.rows_host.numpy()[:roff].view(ROW_DTYPE)
            res.row_start = self.row_start_host.numpy()[:S]
            res.row_count = self.row_count_host.numpy()[:S]
            for ch in self.chunks:               # the chunks' row blocks follow one another in the result buffer
                if ch.row_off:
                    res.row_start[ch.s0:ch.s1] += ch.row_off
        return res


class PipelineResult(object):
    pass
