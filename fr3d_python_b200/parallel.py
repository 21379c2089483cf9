"""Multi-GPU: the path shards by structure (annotator) and by matrix row block (discrepancy);
there is no exchange step, so no collective runs on the data path.  One process per GPU
(torch.distributed, NCCL on GPUs / gloo in the CPU tests); results are gathered to rank 0 once.
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


def annotate_batch_sharded(batch, categories, annotate_fn, rank=None, world_size=None, gather=True):
    """Shard the structures of `batch` over ranks (LPT on nt count), run `annotate_fn(sub_batch,
    categories) -> list of per-structure results` on the local shard and gather the per-structure
    results to rank 0 in the original structure order.  Returns the full list on rank 0, the
    local list elsewhere (or everywhere when gather=False)."""
    dist = _dist()
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
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


def all_vs_all_sharded(centers, rotations, compute_rows, rank=None, world_size=None, block=256):
    """Row-block sharded all-vs-all discrepancy: `compute_rows(row0, row1) -> (row1-row0, N)` array is
    called for this rank's blocks; slabs are gathered to rank 0 and assembled.  Rank 0 returns the
    N x N matrix, others None."""
    dist = _dist()
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
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
