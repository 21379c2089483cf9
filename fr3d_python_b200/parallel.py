"""Multi-GPU: the path shards by structure (annotator) and by matrix row block (discrepancy);
there is no exchange step, so no collective runs on the data path.  One process per GPU
(torch.distributed, NCCL on GPUs / gloo in the CPU tests); results are gathered to rank 0 once,
with tensor collectives (padded row arrays / matrix slabs), never pickled objects.

  lpt_partition, cyclic_row_blocks, triangle_blocks     who owns what
  gather_rows                                           per-rank row arrays -> rank 0 (any backend)
  ShardedAnnotator                                      ONE batch sharded over the ranks (BASELINE configs[3]):
                                                        LPT shard -> engine.Pipeline with the device tail per rank ->
                                                        rows gathered on the device -> one D2H on rank 0
  gather_slabs                                          discrepancy row blocks -> the N x N matrix on rank 0
"""

import numpy as np

from . import packing


def lpt_partition(weights, n_parts):
    """Longest-processing-time bin packing: structures sorted by weight (nt count ~ pair count),
    each placed on the currently lightest rank.  Returns n_parts lists of structure indices, each
    sorted ascending so that results can be merged deterministically."""
    order = sorted(range(len(weights)), key=lambda k: (-weights[k], k))
    loads = [0] * n_parts
    parts = [[] for _ in range(n_parts)]
    for k in order:
        r = min(range(n_parts), key=lambda q: (loads[q], q))
        parts[r].append(k)
        loads[r] += weights[k]
    return [sorted(p) for p in parts]


def cyclic_row_blocks(N, n_parts, block=256):
    """Row blocks of the N x N discrepancy matrix dealt cyclically to ranks (balances the work and
    the output slabs).  Returns per rank a list of (row0, row1)."""
    parts = [[] for _ in range(n_parts)]
    for b, r0 in enumerate(range(0, N, block)):
        parts[b % n_parts].append((r0, min(N, r0 + block)))
    return parts


def _dist():
    import torch.distributed as dist
    return dist


def _rank_world(rank, world_size):
    dist = _dist()
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    return rank, world_size


def gather_rows(rows, row_start, row_count, mine, n_structures, rank=None, world_size=None, scalars=()):
    """Gather every rank's final rows to rank 0 with tensor collectives.

    rows int32 [n_rows, 4] (label, nt1, nt2, crossing), row_start / row_count int32 [len(mine)] for the structures
    `mine` (indices into the whole batch) - torch tensors on the device the process group works on (CUDA for NCCL,
    CPU for gloo).  scalars: extra int64 values summed over ranks (e.g. candidate pairs).
    Returns on rank 0 (rows [total, 4], row_start [n_structures], row_count [n_structures], summed scalars) as
    tensors on that device, with row_start pointing into the concatenated array; None elsewhere.
    Collectives: one all_gather of (n_rows, n_structures, scalars...) - its result sizes the padded buffers -
    then one gather of the rows and one of the per-structure (start, count, structure index) table."""
    import torch
    dist = _dist()
    rank, world_size = _rank_world(rank, world_size)
    dev = rows.device
    mine_t = torch.as_tensor(np.asarray(mine, dtype=np.int32), device=dev)
    if world_size == 1:
        rs = torch.zeros(n_structures, dtype=torch.int32, device=dev)
        rc = torch.zeros(n_structures, dtype=torch.int32, device=dev)
        rs[mine_t.long()] = row_start
        rc[mine_t.long()] = row_count
        return rows, rs, rc, [int(v) for v in scalars]
    head = torch.tensor([int(rows.shape[0]), len(mine)] + [int(v) for v in scalars], dtype=torch.int64, device=dev)
    heads = torch.empty((world_size, head.numel()), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(heads.view(-1), head)
    heads_h = heads.cpu()
    max_rows, max_structs = int(heads_h[:, 0].max()), int(heads_h[:, 1].max())
    rows_pad = torch.zeros((max(max_rows, 1), 4), dtype=torch.int32, device=dev)
    rows_pad[:rows.shape[0]] = rows
    table = torch.zeros((max(max_structs, 1), 3), dtype=torch.int32, device=dev)
    table[:len(mine), 0] = row_start
    table[:len(mine), 1] = row_count
    table[:len(mine), 2] = mine_t
    if rank == 0:
        rows_all = [torch.empty_like(rows_pad) for _ in range(world_size)]
        table_all = [torch.empty_like(table) for _ in range(world_size)]
        dist.gather(rows_pad, rows_all, dst=0)
        dist.gather(table, table_all, dst=0)
    else:
        dist.gather(rows_pad, None, dst=0)
        dist.gather(table, None, dst=0)
        return None
    out_rows, off = [], 0
    rs = torch.zeros(n_structures, dtype=torch.int32, device=dev)
    rc = torch.zeros(n_structures, dtype=torch.int32, device=dev)
    for r in range(world_size):
        nr, ns = int(heads_h[r, 0]), int(heads_h[r, 1])
        out_rows.append(rows_all[r][:nr])
        t = table_all[r][:ns]
        idx = t[:, 2].long()
        rs[idx] = t[:, 0] + off
        rc[idx] = t[:, 1]
        off += nr
    return torch.cat(out_rows), rs, rc, [int(v) for v in heads_h[:, 2:].sum(dim=0)]


HDR_INTS = 64          # per chunk block: 32 int64 of tail counters; [4..11] hold a copy of the search / classify counters


class ShardedAnnotator(object):
    """ONE batch annotated by all ranks (BASELINE configs[3]): structures are dealt to the ranks by LPT on their nt
    count, every rank runs engine.Pipeline (K1..K4 + device tail) on its shard, the final rows are gathered on the
    devices (one NCCL gather of a fixed-size block per rank) and rank 0 copies them to the host once.

    Every rank owns ONE device block laid out per pipeline chunk as [counters | row_start | row_count | rows]; the
    tail kernels write their output straight into it (no packing pass), the block is the send buffer of the gather,
    and rank 0's receive buffer goes to pinned host memory in one copy, where the result object indexes it in place.
    Block sizes are agreed once in the constructor (a calibration pass + one all_reduce(MAX)), so a pass needs no
    host synchronisation before the final copy.  run() returns a tail.BatchAnnotations over the WHOLE batch on
    rank 0 (None elsewhere)."""

    def __init__(self, batch, categories, focused_basepair_cutoffs=None, ideal_hydrogen_bonds=None, rank=None,
                 world_size=None, device=None, n_chunks=2, use_graph=True):
        import torch
        self.use_graph = use_graph
        from .classifiers import NA_pairwise_interactions as NA
        from .engine import Engine, Pipeline, pin_batch
        dist = _dist()
        self.rank, self.world = _rank_world(rank, world_size)
        self.batch = batch
        if not focused_basepair_cutoffs:
            focused_basepair_cutoffs = NA._default_focused(categories.get('basepair', []))
        if not ideal_hydrogen_bonds:
            ideal_hydrogen_bonds = NA.load_ideal_basepair_hydrogen_bonds()
        bt, _atoms, self.labels = NA._tables(focused_basepair_cutoffs, ideal_hydrogen_bonds)
        weights = np.diff(batch.struct_off).astype(int).tolist()
        self.parts = lpt_partition(weights, self.world)
        self.mine = self.parts[self.rank]
        self.sub = packing.select_structures(batch, self.mine)
        self.eng = Engine.get(device)
        self.torch = torch
        dev = self.eng.device
        mask = NA.category_mask(categories)
        self.pinned = pin_batch(torch, self.sub)
        # every rank uses the same number of chunks (the block layout is common); a rank with fewer structures than
        # chunks still has that many (possibly empty) chunk blocks
        K = self.n_chunks = max(1, min(n_chunks, min(len(p) for p in self.parts) or 1))
        # calibration: a pass with self-growing buffers tells the counts of every chunk
        need = np.zeros((K, 4), dtype=np.int64)            # pairs, hits, rows, structures per chunk
        if len(self.mine):
            cal = Pipeline(self.eng, self.sub, self.pinned, bt, mask, n_chunks=K, labels=self.labels, want_hits=False)
            cal.run()
            for k, ch in enumerate(cal.chunks):
                need[k] = (int(ch.counters_host[0]), ch.n_hits, ch.n_rows, ch.s1 - ch.s0)
            del cal
        need_t = torch.as_tensor(need, device=dev)
        if self.world > 1:
            dist.all_reduce(need_t, op=dist.ReduceOp.MAX)
        need = need_t.cpu().numpy()
        self.caps = [(int(p * 1.02) + 1024, int(h * 1.02) + 1024, int(r * 1.02) + 1024) for p, h, r, _ in need]
        self.s_max = [int(x) for x in need[:, 3]]
        # block layout (int32 units), every region a multiple of 4 so that the rows region is 16-byte aligned
        self.off = []
        total = 0
        for k in range(K):
            sm = (self.s_max[k] + 3) // 4 * 4
            hdr, rs, rc, rows = total, total + HDR_INTS, total + HDR_INTS + sm, total + HDR_INTS + 2 * sm
            total = rows + 4 * self.caps[k][2]
            self.off.append((hdr, rs, rc, rows))
        self.block_ints = total
        self.block = torch.zeros(total, dtype=torch.int32, device=dev)
        self.recv = torch.empty((self.world, total), dtype=torch.int32, device=dev) if self.rank == 0 else None
        self.host = torch.empty((self.world, total), dtype=torch.int32).pin_memory() if self.rank == 0 else None

        def row_out(k, n_struct, row_cap):
            hdr, rs, rc, rows = self.off[k]
            b = self.block
            return (b[rows:rows + 4 * self.caps[k][2]].view(torch.uint8), b[rs:rs + max(n_struct, 1)],
                    b[rc:rc + max(n_struct, 1)], b[hdr:hdr + HDR_INTS].view(torch.int64))

        self.pipe = Pipeline(self.eng, self.sub, self.pinned, bt, mask, n_chunks=K, labels=self.labels, want_hits=False,
                             chunk_caps=self.caps, row_out=row_out) if len(self.mine) else None
        if self.pipe is not None and len(self.pipe.chunks) != K:
            raise RuntimeError("chunk count mismatch")
        # where the structures of every rank's chunks sit (the chunking of Pipeline: contiguous, even split)
        self.layout = []
        for r in range(self.world):
            S_r = len(self.parts[r])
            kk = max(1, min(K, S_r)) if S_r else 0
            bounds = [int(round(q * S_r / kk)) for q in range(kk + 1)] if kk else [0]
            self.layout.append(bounds)
        sub_off = {r: np.concatenate([[0], np.cumsum([weights[s] for s in self.parts[r]])]) for r in range(self.world)}
        self.sub_off = sub_off

    def enqueue(self, upload=True):
        """The device part of a pass, asynchronously on the current stream: shard pipeline (upload=False: on the
        records already in HBM), counters into the block headers, NCCL gather of the blocks into rank 0's HBM."""
        torch = self.torch
        dist = _dist()
        if self.pipe is not None:
            def headers():                                 # the search / classify counters travel in the header
                for k, ch in enumerate(self.pipe.chunks):
                    hdr = self.off[k][0]
                    self.block[hdr:hdr + HDR_INTS].view(torch.int64)[4:12].copy_(ch.plan.counters, non_blocking=True)
            if self.use_graph:
                self.pipe.enqueue_graph(upload=upload, extra=headers)      # the rank's whole pass: one graph launch
            else:
                self.pipe.enqueue(upload=upload)
                headers()
        if self.world > 1:
            if self.rank == 0:
                dist.gather(self.block, list(self.recv.unbind(0)), dst=0)
            else:
                dist.gather(self.block, None, dst=0)

    def run(self, upload=True):
        torch, dev = self.torch, self.eng.device
        self.enqueue(upload=upload)
        if self.rank != 0:
            return None
        src = self.recv if self.world > 1 else self.block.view(1, -1)
        self.host.copy_(src, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return self._assemble()

    def _assemble(self):
        from . import _lib
        from .tail import BatchAnnotations, ROW_DTYPE
        host = self.host.numpy()
        S = self.batch.n_structures
        row_start = np.zeros(S, dtype=np.int64)
        row_count = np.zeros(S, dtype=np.int32)
        nt_base = np.zeros(S, dtype=np.int64)
        n_pairs = 0
        for r in range(self.world):
            bounds = self.layout[r]
            mine = self.parts[r]
            for k in range(len(bounds) - 1):
                hdr, rs, rc, rows = self.off[k]
                cnt = host[r, hdr:hdr + HDR_INTS].view(np.int64)
                pairs, hits, nrows = int(cnt[4]), int(cnt[5]), int(cnt[0])
                if int(cnt[6]) != 0 or int(cnt[2]) != 0:
                    raise _lib.Fr3dError("rank %d chunk %d: coordinates or a structure outside the supported ranges (FR3D_E_RANGE)" % (r, k))
                if pairs > self.caps[k][0] or hits > self.caps[k][1] or nrows > self.caps[k][2]:
                    raise _lib.Fr3dError("rank %d chunk %d: capacity exceeded (pairs %d, hits %d, rows %d > %r)"
                                         % (r, k, pairs, hits, nrows, self.caps[k]))
                n_pairs += pairs
                ns = bounds[k + 1] - bounds[k]
                ids = mine[bounds[k]:bounds[k + 1]]
                row_start[ids] = host[r, rs:rs + ns].astype(np.int64) + (r * self.block_ints + rows) // 4
                row_count[ids] = host[r, rc:rc + ns]
                nt_base[ids] = self.sub_off[r][bounds[k]:bounds[k + 1]]
        rows_all = host.reshape(-1).view(ROW_DTYPE)
        return BatchAnnotations(self.batch, rows_all, row_start, row_count, self.labels, n_pairs=n_pairs, nt_base=nt_base)


def annotate_batch_sharded(batch, categories, annotate_fn=None, rank=None, world_size=None, gather=True, **kw):
    """Shard the structures of `batch` over the ranks (LPT on nt count), annotate the local shard, gather to rank 0.

    Default (annotate_fn None): ShardedAnnotator - device pipeline per rank, rows gathered with NCCL; returns a
    tail.BatchAnnotations over the whole batch on rank 0, None elsewhere.
    With annotate_fn(sub_batch, categories) -> list of per-structure results (the host-side protocol, used by the
    CPU tests with gloo): the per-structure results are gathered with gather_object in the original structure order;
    rank 0 returns the full list, the others their local list (everyone the local list when gather=False)."""
    if annotate_fn is None:
        return ShardedAnnotator(batch, categories, rank=rank, world_size=world_size, **kw).run()
    dist = _dist()
    rank, world_size = _rank_world(rank, world_size)
    weights = np.diff(batch.struct_off).astype(int).tolist()
    parts = lpt_partition(weights, world_size)
    mine = parts[rank]
    sub = packing.select_structures(batch, mine)
    local = annotate_fn(sub, categories) if len(mine) else []
    if world_size == 1 or not gather:
        return local
    gathered = [None] * world_size if rank == 0 else None
    dist.gather_object(list(zip(mine, local)), gathered, dst=0)
    if rank != 0:
        return local
    out = [None] * batch.n_structures
    for part in gathered:
        for s, res in part:
            out[s] = res
    return out


# ------------------------------------------------------------------ discrepancy (C5)
def triangle_blocks(N, n_parts, block=256):
    """Ownership of the upper triangle of the N x N discrepancy matrix: row blocks of `block` rows dealt to the ranks
    so that every rank gets the same number of tiles ON OR ABOVE the diagonal (row block b has ~ (nb - b) of them):
    blocks are taken in pairs (b, nb - 1 - b), dealt cyclically.  A rank evaluates only the part of its row blocks
    at or right of the diagonal (fr3d_discrepancy_all_vs_all with col0 = row0); the lower triangle is the mirror
    (FR3D.py:439-441).  Returns per rank a list of (row0, row1)."""
    nb = (N + block - 1) // block
    parts = [[] for _ in range(n_parts)]
    k = 0
    lo, hi = 0, nb - 1
    while lo <= hi:
        r = k % n_parts
        parts[r].append((lo * block, min(N, (lo + 1) * block)))
        if hi != lo:
            parts[r].append((hi * block, min(N, (hi + 1) * block)))
        lo += 1
        hi -= 1
        k += 1
    return [sorted(p) for p in parts]


def all_vs_all_sharded(centers, rotations, compute_rows, rank=None, world_size=None, block=256):
    """Row-block sharded all-vs-all discrepancy (host protocol, used by the CPU tests): `compute_rows(row0, row1) ->
    (row1-row0, N)` array is called for this rank's blocks; slabs are gathered to rank 0 and assembled.  Rank 0
    returns the N x N matrix, others None.  The device form is geometry.discrepancy.all_vs_all_multi_gpu."""
    dist = _dist()
    rank, world_size = _rank_world(rank, world_size)
    N = len(centers)
    blocks = cyclic_row_blocks(N, world_size, block)
    slabs = [(r0, r1, np.asarray(compute_rows(r0, r1))) for r0, r1 in blocks[rank]]
    if world_size == 1:
        gathered = [slabs]
    else:
        gathered = [None] * world_size if rank == 0 else None
        dist.gather_object(slabs, gathered, dst=0)
        if rank != 0:
            return None
    D = np.zeros((N, N))
    for part in gathered:
        for r0, r1, slab in part:
            D[r0:r1] = slab
    return D
