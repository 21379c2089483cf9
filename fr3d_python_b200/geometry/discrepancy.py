"""Drop-in fr3d/geometry/discrepancy.py (matrix_discrepancy :165-233, matrix_discrepancy_cutoff
:235-313) and fr3d/search/discrepancy.py (same arithmetic with default weights), on the CUDA
discrepancy kernels (K6), plus the all-vs-all / one-vs-many forms that replace the Python double
loop of fr3d/search/FR3D.py:433-442 and the scan of fr3d/search/search.py:1073-1085."""

import numpy as np

from .. import _lib
from ..engine import Engine

MAX_POSITIONS = 16


def _pack_instance(centers, rotations):
    """(n,3) centres, (n,9) rotations, (n,) has_rot from the reference's list arguments; an empty
    or None rotation (NA_protein_annotation.py:1705, search/discrepancy.py:196) has no rotation."""
    n = len(centers)
    c = np.zeros((n, 3))
    r = np.zeros((n, 9))
    h = np.zeros(n, dtype=np.uint8)
    for k in range(n):
        c[k] = np.asarray(centers[k], dtype=float).reshape(3)
        rk = rotations[k]
        if rk is not None and np.asarray(rk).shape[0] > 0:
            r[k] = np.asarray(rk, dtype=float).reshape(9)
            h[k] = 1
    return c, r, h


def one_vs_many(q_centers, q_rotations, centers, rotations, cutoff=None, angle_weight=1.0, center_weight=None,
                q_has_rot=None, has_rot=None, device=None):
    """Discrepancy of one query instance against M candidates (arrays (M,n,3), (M,n,3,3)).
    Returns (M,) float64; NaN where matrix_discrepancy_cutoff would return None."""
    eng = Engine.get(device)
    torch = eng.torch
    centers = np.ascontiguousarray(centers, dtype=np.float64)
    M, n = centers.shape[0], centers.shape[1]
    rots = np.ascontiguousarray(rotations, dtype=np.float64).reshape(M, n, 9)
    qc = np.ascontiguousarray(q_centers, dtype=np.float64).reshape(n, 3)
    qr = np.ascontiguousarray(q_rotations, dtype=np.float64).reshape(n, 9)
    with torch.cuda.device(eng.device):
        dev = eng.device
        t = lambda a: torch.from_numpy(a).to(dev)
        dqc, dqr, dc, dr = t(qc), t(qr), t(centers), t(rots)
        dqh = t(np.ascontiguousarray(q_has_rot, dtype=np.uint8)) if q_has_rot is not None else None
        dh = t(np.ascontiguousarray(has_rot, dtype=np.uint8)) if has_rot is not None else None
        dw = t(np.ascontiguousarray(center_weight, dtype=np.float64)) if center_weight is not None else None
        out = torch.empty(M, dtype=torch.float64, device=dev)
        _lib.check(eng.lib.fr3d_discrepancy_one_vs_many(
            dqc.data_ptr(), dqr.data_ptr(), dqh.data_ptr() if dqh is not None else None, dc.data_ptr(),
            dr.data_ptr(), dh.data_ptr() if dh is not None else None, M, n,
            -1.0 if cutoff is None else float(cutoff), float(angle_weight),
            dw.data_ptr() if dw is not None else None, out.data_ptr(), eng.stream_ptr()),
            "fr3d_discrepancy_one_vs_many")
        eng.launches += 1
        return out.cpu().numpy()


def all_vs_all(centers, rotations, has_rot=None, angle_weight=1.0, rows=None, device=None, out_device=False):
    """All-vs-all discrepancy matrix of N instances with n positions each: centers (N,n,3),
    rotations (N,n,3,3).  rows=(row0,row1) computes a row block (multi-GPU sharding).  The diagonal
    is 0 like the reference's heat-map loop (FR3D.py:433-442)."""
    eng = Engine.get(device)
    torch = eng.torch
    centers = np.ascontiguousarray(centers, dtype=np.float64)
    N, n = centers.shape[0], centers.shape[1]
    rots = np.ascontiguousarray(rotations, dtype=np.float64).reshape(N, n, 9)
    row0, row1 = (0, N) if rows is None else rows
    with torch.cuda.device(eng.device):
        dc = torch.from_numpy(centers).to(eng.device)
        dr = torch.from_numpy(rots).to(eng.device)
        dh = torch.from_numpy(np.ascontiguousarray(has_rot, dtype=np.uint8)).to(eng.device) if has_rot is not None else None
        out = torch.empty((row1 - row0, N), dtype=torch.float64, device=eng.device)
        _lib.check(eng.lib.fr3d_discrepancy_all_vs_all(dc.data_ptr(), dr.data_ptr(),
                                                       dh.data_ptr() if dh is not None else None, N, n,
                                                       float(angle_weight), row0, row1, out.data_ptr(),
                                                       eng.stream_ptr()), "fr3d_discrepancy_all_vs_all")
        eng.launches += 1
        return out if out_device else out.cpu().numpy()


def matrix_discrepancy(centers1, rotations1, centers2, rotations2, angle_weight=None, center_weight=None):
    """discrepancy.py:165-233."""
    n = len(centers1)
    assert len(centers2) == n
    assert len(rotations1) == n
    assert len(rotations2) == n
    assert n >= 2
    if n > MAX_POSITIONS:
        raise _lib.Fr3dError("matrix_discrepancy: at most %d positions are supported on the GPU" % MAX_POSITIONS)
    if not angle_weight:
        angle_weight = 1.0
    if isinstance(angle_weight, (list, tuple, np.ndarray)):      # search copy: list, only [0] used for n == 2
        angle_weight = float(angle_weight[0]) if n == 2 else 1.0
    c1, r1, h1 = _pack_instance(centers1, rotations1)
    c2, r2, h2 = _pack_instance(centers2, rotations2)
    cw = None
    if center_weight is not None and len(center_weight) == n:
        cw = np.asarray(center_weight, dtype=float)
    d = one_vs_many(c1, r1, c2[None], r2[None], None, angle_weight, cw, h1, h2[None])
    return float(d[0])


def matrix_discrepancy_cutoff(centers1, rotations1, centers2, rotations2, cutoff, angle_weight=None,
                              center_weight=None):
    """discrepancy.py:235-313 -> float, or None once the running error exceeds (n*cutoff)^2."""
    n = len(centers1)
    assert len(centers2) == n
    assert len(rotations1) == n
    assert len(rotations2) == n
    assert n >= 2
    if n > MAX_POSITIONS:
        raise _lib.Fr3dError("matrix_discrepancy_cutoff: at most %d positions are supported" % MAX_POSITIONS)
    aw = 1.0
    if angle_weight is not None and n == 2:
        aw = float(angle_weight[0]) if isinstance(angle_weight, (list, tuple, np.ndarray)) else float(angle_weight)
    c1, r1, h1 = _pack_instance(centers1, rotations1)
    c2, r2, h2 = _pack_instance(centers2, rotations2)
    cw = None
    if center_weight is not None and len(center_weight) == n:
        cw = np.asarray(center_weight, dtype=float)
    d = one_vs_many(c1, r1, c2[None], r2[None], None if n == 2 else cutoff, aw, cw, h1, h2[None])
    return None if np.isnan(d[0]) else float(d[0])


def discrepancy(ntlist1, ntlist2, centers=['base'], base_weights=1.0, P_weights=1.0, C1star_weights=1.0,
                angleweight=1.0):
    """discrepancy.py:38-162 for the default arguments (base centres, unit weights), where it equals
    matrix_discrepancy of the components' centres and rotation matrices (the reference
    implementation itself no longer runs on Python 3: xrange, :90)."""
    assert len(ntlist1) == len(ntlist2)
    if list(centers) != ['base']:
        raise NotImplementedError("only the default centers=['base'] form is accelerated")
    c1 = [nt.centers['base'] for nt in ntlist1]
    r1 = [np.asarray(nt.rotation_matrix) for nt in ntlist1]
    c2 = [nt.centers['base'] for nt in ntlist2]
    r2 = [np.asarray(nt.rotation_matrix) for nt in ntlist2]
    return matrix_discrepancy(c1, r1, c2, r2, angle_weight=angleweight)


def fp64_peak_tflops(device=None, iters=20000, reps=5):
    """Measured DFMA throughput of the device (fr3d_fp64_fma_probe timed with CUDA events; best of reps), in
    TFLOP/s with an FMA counted as two flops.  The roofline denominator of K5 / K6."""
    import ctypes as C
    eng = Engine.get(device)
    torch = eng.torch
    with torch.cuda.device(eng.device):
        sink = torch.zeros(8, dtype=torch.float64, device=eng.device)
        flops = C.c_int64(0)
        best = 0.0
        for _ in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(eng.lib.fr3d_fp64_fma_probe(int(iters), sink.data_ptr(), C.byref(flops), eng.stream_ptr()),
                       "fr3d_fp64_fma_probe")
            e1.record()
            torch.cuda.synchronize()
            best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        eng.launches += reps + 1
    return best


class AllVsAllSharded(object):
    """The N x N discrepancy matrix on several GPUs (BASELINE configs[4]): the instances are replicated,
    the UPPER triangle is dealt to the ranks in row blocks with equal tile counts (parallel.triangle_blocks), every
    rank evaluates only its blocks' part at or right of the diagonal (fr3d_discrepancy_rows_upper) into one packed
    buffer, the buffers are gathered to rank 0 with one NCCL gather (padded to the largest share), and rank 0
    places and mirrors the blocks into the matrix (fr3d_discrepancy_assemble).  No exchange between the ranks.
    run() returns the matrix as a DEVICE tensor on rank 0 (None elsewhere); to_host() copies it out."""

    def __init__(self, centers, rotations, has_rot=None, angle_weight=1.0, rank=0, world_size=1, device=None, block=256):
        from .. import parallel
        self.eng = Engine.get(device)
        torch = self.torch = self.eng.torch
        dev = self.eng.device
        self.rank, self.world = rank, world_size
        centers = np.ascontiguousarray(centers, dtype=np.float64)
        self.N, self.n = centers.shape[0], centers.shape[1]
        rots = np.ascontiguousarray(rotations, dtype=np.float64).reshape(self.N, self.n, 9)
        self.aw = float(angle_weight)
        self.host_in = (torch.from_numpy(centers).pin_memory(), torch.from_numpy(rots).pin_memory())
        self.dc = torch.empty_like(self.host_in[0], device=dev)
        self.dr = torch.empty_like(self.host_in[1], device=dev)
        self.dh = torch.from_numpy(np.ascontiguousarray(has_rot, dtype=np.uint8)).to(dev) if has_rot is not None else None
        self.parts = parallel.triangle_blocks(self.N, world_size, block)
        self.offsets = []          # per rank: [(row0, row1, offset in doubles)]
        sizes = []
        for part in self.parts:
            off, lst = 0, []
            for r0, r1 in part:
                lst.append((r0, r1, off))
                off += (r1 - r0) * (self.N - r0)
            self.offsets.append(lst)
            sizes.append(off)
        self.share = max(sizes + [1])
        self.buf = torch.empty(self.share, dtype=torch.float64, device=dev)
        self.recv = torch.empty((world_size, self.share), dtype=torch.float64, device=dev) if (rank == 0 and world_size > 1) else None
        self.D = torch.zeros((self.N, self.N), dtype=torch.float64, device=dev) if rank == 0 else None
        self.host_out = None
        self.in_bytes = centers.nbytes + rots.nbytes

    def upload(self):
        self.dc.copy_(self.host_in[0], non_blocking=True)
        self.dr.copy_(self.host_in[1], non_blocking=True)

    def run(self, upload=False):
        eng, torch = self.eng, self.torch
        import torch.distributed as dist
        with torch.cuda.device(eng.device):
            if upload:
                self.upload()
            st = eng.stream_ptr()
            hp = self.dh.data_ptr() if self.dh is not None else None
            if self.world == 1:          # one GPU: tiles on or above the diagonal, mirrored inside the kernel
                self.evaluate_share()
                return self.D
            for r0, r1, off in self.offsets[self.rank]:
                _lib.check(eng.lib.fr3d_discrepancy_rows_upper(self.dc.data_ptr(), self.dr.data_ptr(), hp, self.N, self.n,
                                                               self.aw, r0, r1, self.buf[off:].data_ptr(), st),
                           "fr3d_discrepancy_rows_upper")
                eng.launches += 1
            if self.world > 1:
                if self.rank == 0:
                    dist.gather(self.buf, list(self.recv.unbind(0)), dst=0)
                else:
                    dist.gather(self.buf, None, dst=0)
                    return None
            for r in range(self.world):
                src = self.recv[r] if self.world > 1 else self.buf
                for r0, r1, off in self.offsets[r]:
                    _lib.check(eng.lib.fr3d_discrepancy_assemble(src[off:].data_ptr(), self.N, r0, r1, self.D.data_ptr(), st),
                               "fr3d_discrepancy_assemble")
                    eng.launches += 1
        return self.D

    def evaluate_share(self):
        """Only the evaluation launches of this rank (no gather, no assembly): what the roofline figure times."""
        eng = self.eng
        st = eng.stream_ptr()
        hp = self.dh.data_ptr() if self.dh is not None else None
        if self.world == 1:
            _lib.check(eng.lib.fr3d_discrepancy_all_vs_all(self.dc.data_ptr(), self.dr.data_ptr(), hp, self.N, self.n,
                                                           self.aw, 0, self.N, self.D.data_ptr(), st),
                       "fr3d_discrepancy_all_vs_all")
            eng.launches += 1
            return
        for r0, r1, off in self.offsets[self.rank]:
            _lib.check(eng.lib.fr3d_discrepancy_rows_upper(self.dc.data_ptr(), self.dr.data_ptr(), hp, self.N, self.n,
                                                           self.aw, r0, r1, self.buf[off:].data_ptr(), st),
                       "fr3d_discrepancy_rows_upper")
            eng.launches += 1

    def to_host(self):
        torch = self.torch
        if self.host_out is None:
            self.host_out = torch.empty((self.N, self.N), dtype=torch.float64).pin_memory()
        self.host_out.copy_(self.D, non_blocking=True)
        torch.cuda.current_stream(self.eng.device).synchronize()
        return self.host_out.numpy()


def all_vs_all_multi_gpu(centers, rotations, has_rot=None, angle_weight=1.0, device=None):
    """all_vs_all() over the ranks of the initialised torch.distributed group (one process per GPU): rank 0 gets
    the N x N matrix (numpy), the others None.  Works without a group as a single rank."""
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    sh = AllVsAllSharded(centers, rotations, has_rot, angle_weight, rank, world, device)
    D = sh.run(upload=True)
    return sh.to_host() if D is not None else None
