"""Device side of the annotator: owns torch buffers, calls the C ABI (K1-K4).

PyTorch is used only for device memory, pinned host memory and streams.
Everything fails loudly when CUDA or the extension is unavailable — there is
no CPU path in this package.
"""

import ctypes as C

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
        t = {}
        for name in ("center", "rot", "base_xyz", "bb_xyz", "base_mask", "bb_mask", "meta"):
            src = pinned[name] if pinned is not None else torch.from_numpy(_as_signed(getattr(batch, name)))
            if name == "base_xyz" and src.shape[1] < _lib.BASE_SLOTS:       # compact pinned form, see pin_batch
                t[name] = torch.zeros((batch.n, _lib.BASE_SLOTS, 3), dtype=src.dtype, device=device)
                t[name][:, :src.shape[1]].copy_(src.to(device, non_blocking=True))
            else:
                t[name] = src.to(device, non_blocking=True)
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
    def fit_frames(self, dbatch, place_h=True):
        torch = self.torch
        status = torch.empty(dbatch.n, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.fr3d_frames_fit(C.byref(dbatch.nts), status.data_ptr(), 1 if place_h else 0,
                                                self.stream_ptr()), "fr3d_frames_fit")
        self.launches += 1
        return status

    # ---------------------------------------------------------------- plans
    def make_plan(self, n_nt, pair_cap=None, hit_cap=None):
        """Pre-allocate every device buffer one pass over an n_nt batch needs, so that the pass
        itself is launches only (no allocation, no host sync)."""
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
        p.status = torch.empty(max(n_nt, 1), dtype=torch.int32, device=dev)
        return p

    def run_plan(self, dbatch, plan, boxtable, category_mask, fit_frames=True, events=None, index_offset=0):
        """K1 -> K2/K3 -> K4 on the current stream, asynchronously.  `events` (dict) receives
        torch CUDA events around the classify launch for per-kernel timing."""
        torch = self.torch
        bx, ix = boxtable.device(torch, self.device)
        st = self.stream_ptr()
        with torch.cuda.device(self.device):
            if fit_frames:
                _lib.check(self.lib.fr3d_frames_fit(C.byref(dbatch.nts), plan.status.data_ptr(), 1, st),
                           "fr3d_frames_fit")
                self.launches += 1
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

    def read_counts(self, plan):
        """(n_pairs, n_hits) after a sync; raises on range errors; None if a capacity was exceeded."""
        cnt = plan.counters.cpu()
        if int(cnt[2]) != 0:
            raise _lib.Fr3dError("fr3d_pair_search: base centre outside the spatial-hash range (FR3D_E_RANGE)")
        return int(cnt[0]), int(cnt[1])

    # ---------------------------------------------------------------- whole path
    def annotate_batch(self, batch, boxtable, category_mask, pinned=None, return_frames=False, plan=None):
        """H2D -> [K1] -> K2/K3 -> K4 -> D2H.  Returns (hits structured array, n_pairs, extras).
        Capacity overflow (FR3D_E_CAPACITY) re-runs with the exact counts; nothing is truncated."""
        torch = self.torch
        with torch.cuda.device(self.device):
            db = DeviceBatch(torch, self.device, batch, pinned)
            if plan is None or plan.n_nt != batch.n:
                plan = self.make_plan(batch.n)
            fit = not batch.frames_given
            while True:
                self.run_plan(db, plan, boxtable, category_mask, fit_frames=fit)
                n_pairs, n_hits = self.read_counts(plan)
                if n_pairs <= plan.pair_cap and n_hits <= plan.hit_cap:
                    break
                fit = False          # frames are already in place in the device batch
                plan = self.make_plan(batch.n, max(n_pairs, plan.pair_cap) + 1024,
                                      max(n_hits, 2 * max(n_pairs, 1)) + 1024)
            hits = plan.hits[: n_hits * HIT_DTYPE.itemsize].cpu().numpy().view(HIT_DTYPE)
            extras = {"plan": plan}
            if return_frames:
                extras["center"] = db.t["center"].cpu().numpy()
                extras["rot"] = db.t["rot"].cpu().numpy()
                extras["base_xyz"] = db.t["base_xyz"].cpu().numpy()
                extras["base_mask"] = db.t["base_mask"].cpu().numpy().view(np.uint32)
                extras["status"] = plan.status[:batch.n].cpu().numpy() if not batch.frames_given else None
                extras["pairs"] = (plan.pair_i[:n_pairs].cpu().numpy(), plan.pair_j[:n_pairs].cpu().numpy(),
                                   plan.pair_rank[:n_pairs].cpu().numpy())
        return hits, n_pairs, extras


class Plan(object):
    pass


_FIELDS = ("center", "rot", "base_xyz", "bb_xyz", "base_mask", "bb_mask", "meta")


class Pipeline(object):
    """Chunked, multi-stream form of Engine.annotate_batch for large batches: structures are
    independent, so the batch is cut into contiguous chunks of whole structures and chunk k+1's
    H2D copy overlaps chunk k's kernels and chunk k-1's D2H copy.  All device / pinned buffers are
    allocated once and re-used across calls.  Frames are always fitted on the device (K1), so the
    centre / rotation planes are not uploaded.  The input is fixed at construction: each chunk's planes are laid out
    back to back in one pinned blob here, and run() uploads those blobs (build a new Pipeline for new coordinates)."""

    def __init__(self, engine, batch, pinned, boxtable, category_mask, n_chunks=4, pair_cap=None, hit_cap=None,
                 n_streams=3):
        torch = engine.torch
        self.eng, self.batch, self.pinned, self.bt, self.mask = engine, batch, pinned, boxtable, category_mask
        if batch.frames_given:
            raise _lib.Fr3dError("Pipeline uploads observed atoms only; frames are fitted on the device")
        S = batch.n_structures
        n_chunks = max(1, min(n_chunks, S))
        bounds = [int(round(k * S / n_chunks)) for k in range(n_chunks + 1)]
        self.chunks = []
        dev = engine.device
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(min(n_streams, n_chunks))]
        for k in range(n_chunks):
            lo, hi = int(batch.struct_off[bounds[k]]), int(batch.struct_off[bounds[k + 1]])
            n = hi - lo
            ch = Plan()
            ch.lo, ch.hi, ch.n = lo, hi, n
            ch.stream = self.streams[k % len(self.streams)]
            # One host -> device copy per chunk: the chunk's slices of the uploaded planes sit back to back in ONE
            # pinned blob (built here, once) and land in ONE device blob whose sub-ranges ARE the device planes
            # (PCIe moves 40 MB pieces at ~55 GB/s but 2-10 MB pieces at ~36 GB/s, profiles/pcie_probe.json).  The
            # compact base plane (slots 0..10) is widened to 16 slots on the device.
            compact = pinned["base_xyz"].shape[1] < _lib.BASE_SLOTS
            uploaded = [name for name in _FIELDS if name not in ("center", "rot")]
            offs, total = {}, 0
            for name in uploaded:
                src = pinned[name][lo:hi]
                offs[name] = (total, src.numel() * src.element_size(), src.dtype, tuple(src.shape))
                total += (offs[name][1] + 255) // 256 * 256
            ch.host_blob = torch.empty(max(total, 1), dtype=torch.uint8).pin_memory()
            ch.dev_blob = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
            ch.dev = {}
            ch.base_stage = None
            for name in uploaded:
                o, nb, dt, shape = offs[name]
                ch.host_blob[o:o + nb].view(dt).view(shape).copy_(pinned[name][lo:hi])
                view = ch.dev_blob[o:o + nb].view(dt).view(shape)
                if name == "base_xyz" and compact:
                    ch.base_stage = view
                    ch.dev[name] = torch.zeros((n, _lib.BASE_SLOTS, 3), dtype=dt, device=dev)
                else:
                    ch.dev[name] = view
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
            ch.plan = engine.make_plan(n, pc, hc)
            ch.counters_host = torch.zeros(8, dtype=torch.int64).pin_memory()
            ch.ev = torch.cuda.Event()
            self.chunks.append(ch)
        total_hc = sum(ch.plan.hit_cap for ch in self.chunks)
        self.hits_host = torch.empty(total_hc * HIT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
        self.hits_np = self.hits_host.numpy()
        self.h2d_bytes = sum(int(pinned[name][0:1].element_size()) * int(np.prod(pinned[name].shape))
                             for name in _FIELDS if name not in ("center", "rot"))

    def run(self):
        """One pass.  Returns (hits, n_pairs): hits is a structured-array VIEW of the pipeline's pinned
        result buffer with GLOBAL nt indices (valid until the next run())."""
        eng, torch = self.eng, self.eng.torch
        main = torch.cuda.current_stream(eng.device)
        start = torch.cuda.Event()
        start.record(main)
        for ch in self.chunks:
            ch.stream.wait_event(start)
            with torch.cuda.stream(ch.stream):
                ch.dev_blob.copy_(ch.host_blob, non_blocking=True)
                if ch.base_stage is not None:
                    ch.dev["base_xyz"][:, :ch.base_stage.shape[1]].copy_(ch.base_stage, non_blocking=True)
                eng.run_plan(ch, ch.plan, self.bt, self.mask, fit_frames=True, index_offset=ch.lo)
                ch.counters_host.copy_(ch.plan.counters, non_blocking=True)
                ch.ev.record(ch.stream)
        n_pairs = 0
        off = 0
        for ch in self.chunks:
            ch.ev.synchronize()
            cnt = ch.counters_host
            if int(cnt[2]) != 0:
                raise _lib.Fr3dError("fr3d_pair_search: base centre outside the spatial-hash range (FR3D_E_RANGE)")
            npairs, nhits = int(cnt[0]), int(cnt[1])
            if npairs > ch.plan.pair_cap or nhits > ch.plan.hit_cap:
                raise _lib.Fr3dError("pipeline chunk capacity exceeded (pairs %d/%d, hits %d/%d): rebuild the "
                                     "Pipeline with larger capacities" % (npairs, ch.plan.pair_cap, nhits,
                                                                          ch.plan.hit_cap))
            n_pairs += npairs
            nb = nhits * HIT_DTYPE.itemsize
            with torch.cuda.stream(ch.stream):
                self.hits_host[off:off + nb].copy_(ch.plan.hits[:nb], non_blocking=True)
            off += nb
        for st in self.streams:
            st.synchronize()
        return self.hits_np[:off].view(HIT_DTYPE), n_pairs
