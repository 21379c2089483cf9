"""The N > 1 path on CPU: two gloo ranks shard a batch by structure (LPT), annotate their shards
(the CUDA stage is replaced by the oracle here: this test is about the sharding / gathering logic)
and gather the per-structure results to rank 0; the discrepancy matrix is sharded by cyclic row
blocks the same way."""

import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as tmp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from fr3d_python_b200 import parallel, synthetic  # noqa: E402

CATS = {'basepair': ['cWW', 'tHS', 'tSH'], 'stacking': []}


def _oracle_annotate(sub, categories):
    from oracle import fr3d_oracle as O
    out = []
    for s in range(sub.n_structures):
        spec = synthetic.batch_to_spec(sub, s)
        nts = O.build_nts(spec)
        res, c2i, _ = O.annotate(spec, categories, nts)
        out.append(({k: [(nts[a].unit_id(), nts[b].unit_id(), c) for a, b, c in v] for k, v in res.items()},
                    {k: set(v) for k, v in c2i.items()}))
    return out


def _small_batch():
    spec = synthetic.load_golden_spec("spec_1s72_strand")
    base = synthetic.packing.pack_specs(spec)
    return synthetic.make_structure_batch(base, 5, seed0=100, sigma=0.05)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch = _small_batch()
    res = parallel.annotate_batch_sharded(batch, CATS, _oracle_annotate)
    rng = np.random.default_rng(3)
    C = rng.normal(size=(70, 4, 3))

    def rows(r0, r1):      # stand-in for the CUDA row-block kernel
        return np.linalg.norm(C[r0:r1, None] - C[None], axis=(2, 3))

    D = parallel.all_vs_all_sharded(C, None, rows, block=16)
    # the tensor-collective gather of final rows (what ShardedAnnotator runs over NCCL): every rank holds the rows
    # of its LPT shard of 7 fake structures, structure s having s + 1 rows, stored in reverse structure order
    import torch
    parts = parallel.lpt_partition([3, 1, 4, 1, 5, 9, 2], world)
    mine = parts[rank]
    blocks, start, count, off = [], [], [], 0
    for s in reversed(mine):
        blocks.append(np.stack([np.full(s + 1, 100 + s), np.arange(s + 1), np.arange(s + 1) * 2, np.full(s + 1, s)], axis=1))
    rows_t = torch.as_tensor(np.concatenate(blocks).astype(np.int32))
    pos = {}
    for s in reversed(mine):
        pos[s] = off
        off += s + 1
    got = parallel.gather_rows(rows_t, torch.as_tensor(np.array([pos[s] for s in mine], dtype=np.int32)),
                               torch.as_tensor(np.array([s + 1 for s in mine], dtype=np.int32)), mine, 7, rank, world,
                               scalars=(10 * (rank + 1),))
    if rank == 0:
        q.put((res, D, [t.numpy() if hasattr(t, "numpy") else t for t in got]))
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res, D, gathered = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    batch = _small_batch()
    want = _oracle_annotate(batch, CATS)
    assert len(res) == 5
    for got, exp in zip(res, want):
        assert got[0] == exp[0] and got[1] == exp[1]
    rng = np.random.default_rng(3)
    C = rng.normal(size=(70, 4, 3))
    assert np.array_equal(D, np.linalg.norm(C[:, None] - C[None], axis=(2, 3)))
    rows, rs, rc, scalars = gathered
    assert scalars == [30] and len(rows) == sum(range(1, 8))
    for s in range(7):
        blk = rows[rs[s]:rs[s] + rc[s]]
        assert rc[s] == s + 1 and np.all(blk[:, 0] == 100 + s) and np.array_equal(blk[:, 1], np.arange(s + 1))


def test_partition_helpers():
    parts = parallel.lpt_partition([316] * 10 + [2844], 4)
    assert sorted(sum(parts, [])) == list(range(11))
    loads = [sum(([316] * 10 + [2844])[k] for k in p) for p in parts]
    assert max(loads) == 2844 and min(loads) >= 948
    blocks = parallel.cyclic_row_blocks(1000, 3, 256)
    covered = sorted(b for part in blocks for b in part)
    assert covered == [(0, 256), (256, 512), (512, 768), (768, 1000)]
    # upper-triangle ownership: every row block once, tiles on or above the diagonal balanced across ranks
    for N, P in ((20000, 8), (20000, 4), (1000, 3), (300, 2)):
        tri = parallel.triangle_blocks(N, P, 256)
        assert sorted(b for part in tri for b in part) == [(r, min(N, r + 256)) for r in range(0, N, 256)]
        work = [sum((r1 - r0) * (N - r0) for r0, r1 in part) for part in tri]
        assert max(work) <= 1.08 * (sum(work) / P) + 256 * 256 * 2
