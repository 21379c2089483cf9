"""bench.py --workload c5: all-vs-all geometric discrepancy of 20 000 motif instances, n = 7 positions (BASELINE
configs[4]; instance k = the sarcin-ricin 7-nt frame set under a random rigid motion + noise, seed
5000).  A step = the whole N x N matrix: every unordered instance pair evaluated once (Kabsch fit + per-position
rotation angles), both triangles filled.

  value   instances resident in HBM -> matrix assembled in rank 0's HBM (N > 1: the upper triangle dealt to the ranks
          in row blocks, blocks gathered with one NCCL gather, mirrored on rank 0); unit: instance pairs / s
  e2e     pinned host instance arrays -> H2D -> ... -> the matrix in rank 0's pinned host memory (3.2 GB D2H)
  roofline: fp64, 154 n + 600 flop per pair against the DFMA peak MEASURED in this run
  cpu_baseline / --impl reference: the unmodified reference's matrix_discrepancy in the double loop of
  fr3d/search/FR3D.py:433-442 on a block of instances, one row range per worker
"""

import gzip
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
METRIC = "instance_pairs_discrepancy_per_sec"
_CPU = {}


def instances(N):
    from fr3d_python_b200 import synthetic
    with gzip.open(os.path.join(ROOT, "tests", "golden", "geometry.json.gz"), "rt") as f:
        av = json.load(f)["all_vs_all"]
    return synthetic.make_instances(av["base_centers"], av["base_rotations"], N, 5000)


def config(N, n, world):
    return {"workload": "C5: all-vs-all geometric discrepancy, %d instances x %d positions (sarcin-ricin frames, random "
                        "rigid motion, centres + N(0, 0.5 A), rotations x N(0, 0.15 rad), seed 5000)" % (N, n),
            "instances": N, "positions": n, "unordered_pairs": N * (N - 1) // 2,
            "l2_policy": "13 MB of instance data stays in L2 by design; the 3.2 GB output is larger than L2",
            "parallelism": "upper triangle in row blocks of 256 dealt to the ranks (equal tile counts), one NCCL gather, "
                           "mirrored on rank 0" if world > 1 else "one GPU: tiles on or above the diagonal, mirrored in the kernel"}


def _cpu_rows(rows):
    import bench
    C, R = _CPU["C"], _CPU["R"]
    M = len(C)
    if bench.reference_available():
        sys.dont_write_bytecode = True
        if bench.REF_DIR not in sys.path:
            sys.path.insert(0, bench.REF_DIR)
        from fr3d.geometry.discrepancy import matrix_discrepancy
    else:
        from oracle.fr3d_oracle import matrix_discrepancy
    out = 0.0
    for i in rows:
        for j in range(i + 1, M):                 # FR3D.py:433-442
            out += matrix_discrepancy(list(C[i]), list(R[i]), list(C[j]), list(R[j]))
    return out


def cpu_block(pool, M, procs):
    rows = list(range(M))
    jobs = [rows[k::procs * 4] for k in range(procs * 4)]         # interleaved: row i has M - 1 - i pairs
    t0 = time.perf_counter()
    pool.map(_cpu_rows, jobs, chunksize=1)
    return M * (M - 1) // 2, time.perf_counter() - t0


def cpu_setup(M):
    C, R = instances(M)
    _CPU["C"], _CPU["R"] = C, R


def run_reference_arm(args):
    import bench
    cores = os.cpu_count() or 1
    procs = min(cores, 64)
    M = max(64, int(np.sqrt(2 * procs * 8000)))          # ~8 000 pairs (~2 s) per worker and step
    cpu_setup(M)
    with mp.get_context("fork").Pool(procs) as pool:
        for _ in range(args.warmup):
            cpu_block(pool, M, procs)
        pairs_total, t_total = 0, 0.0
        for _ in range(args.steps):
            p, dt = cpu_block(pool, M, procs)
            pairs_total += p
            t_total += dt
    value = pairs_total / t_total
    n = _CPU["C"].shape[1]
    what = "unmodified fr3d.geometry.discrepancy.matrix_discrepancy" if bench.reference_available() else "oracle port"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * t_total / max(args.steps, 1), "higher_is_better": True,
            "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config(args.instances, n, args.gpus),
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": procs, "kind": bench.cpu_kind(),
                             "sample": "the first %d instances (%d pairs) per step, %s in the FR3D.py:433-442 double loop"
                                       % (M, M * (M - 1) // 2, what)},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def time_ordering(torch, D, N, with_cpu):
    """The step that follows the matrix in fr3d/search/FR3D.py:446-467: treePenalizedPathLength(matrix, 100, 59) of the
    search copy + reorderSymmetricMatrix, on leading blocks of the matrix that is already in HBM (K8).  The reference
    caps this at 300 candidates (MAXCANDIDATESHEATMAP); its algorithm (oracle port, pinned) is timed at that size."""
    from fr3d_python_b200.search import orderBySimilarity as sobs
    out = {"what": "search/orderBySimilarity.treePenalizedPathLength(block, 100, 59) + reorderSymmetricMatrix on the leading "
                   "n x n block of the resident matrix (tree penalty by average linkage, 100 greedy insertions, one CTA each)",
           "sizes": {}}
    for n in (300, 1000, 3000):
        if n > N:
            break
        block = D[:n, :n].contiguous()
        sobs.treePenalizedPathLength(block, 100, 59)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        order = sobs.treePenalizedPathLength(block, 100, 59)
        sobs.reorderSymmetricMatrix(block, order)
        torch.cuda.synchronize()
        out["sizes"][str(n)] = {"ms": 1e3 * (time.perf_counter() - t0)}
        if n == 300 and with_cpu:
            from oracle import ordering_oracle as oo
            host = block.cpu().numpy()
            t0 = time.perf_counter()
            ref = oo.treePenalizedPathLength_search(host, 100, 59)
            out["sizes"]["300"].update({"cpu_port_ms": 1e3 * (time.perf_counter() - t0), "equal_to_cpu_port": bool(ref == order)})
    return out


def main(args):
    import bench
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    N = args.instances
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        procs = min(cores, 64)
        M = max(64, int(np.sqrt(2 * procs * 20000)))
        cpu_setup(M)
        with mp.get_context("fork").Pool(procs) as pool:
            pairs, dt = cpu_block(pool, M, procs)
        what = "unmodified fr3d.geometry.discrepancy.matrix_discrepancy" if bench.reference_available() else "oracle port"
        cpu_baseline = {"value": pairs / dt, "unit": "pairs/s", "cores": procs, "kind": bench.cpu_kind(),
                        "sample": "the first %d instances (%d pairs), %s in the FR3D.py:433-442 double loop, %d-process "
                                  "pool, %.1f s" % (M, pairs, what, procs, dt)}
    if world > 1:
        bench.bind_to_gpu_numa_node(local_rank)
    import torch
    import torch.distributed as dist
    from fr3d_python_b200.engine import Engine
    from fr3d_python_b200.geometry import discrepancy as disc
    torch.cuda.set_device(local_rank)
    if world > 1:
        bench.init_distributed(torch, dist, local_rank)
    sampler = bench.ClockSampler(local_rank) if rank == 0 else None
    eng = Engine.get(local_rank)
    C, R = instances(N)
    n = C.shape[1]
    sh = disc.AllVsAllSharded(C, R, rank=rank, world_size=world, device=local_rank)
    sh.upload()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        sh.run()
    barrier()
    peak = disc.fp64_peak_tflops(local_rank)
    launches0 = eng.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        evs[k][0].record()
        sh.run()
        evs[k][1].record()
    barrier()
    ms_total = float(sum(a.elapsed_time(b) for a, b in evs))
    launches = eng.launches - launches0
    # kernel alone on this rank's share (roofline numerator): the evaluation launches without gather / assemble
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for k in range(args.steps):
        sh.evaluate_share()
    e1.record()
    barrier()
    kernel_ms = e0.elapsed_time(e1) / args.steps
    # unordered pairs this rank is responsible for (its row blocks' part right of the diagonal)
    my_pairs = sum((r1 - r0) * (N - r0) - (r1 - r0) * (r1 - r0 + 1) // 2 for r0, r1, off in sh.offsets[rank])
    # e2e: host arrays in, host matrix out
    e2e_steps = max(1, min(args.steps, 3))
    sh.run(upload=True)
    if rank == 0:
        sh.to_host()
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        D = sh.run(upload=True)
        if rank == 0:
            Dh = sh.to_host()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    barrier()
    chk = None
    if rank == 0:
        sub = Dh[:64, :64]
        chk = {"symmetric": bool(np.array_equal(Dh[:512, :512], Dh[:512, :512].T)), "diagonal_zero": bool(np.all(np.diag(sub) == 0)),
               "sum_first_1024x1024": float(Dh[:1024, :1024].sum())}
    ordering = None
    if rank == 0 and world == 1 and D is not None:
        ordering = time_ordering(torch, D, N, not args.no_cpu_baseline)
    clocks = sampler.stop() if sampler is not None else None
    vals = torch.tensor([ms_total, t_e2e, kernel_ms], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    ms_total, t_e2e, kernel_ms_max = vals.tolist()
    if rank == 0:
        pairs = N * (N - 1) // 2
        flop_pair = 154 * n + 600
        achieved = my_pairs * flop_pair / (kernel_ms * 1e-3) / 1e12
        line = {"metric": METRIC, "value": pairs * args.steps / (ms_total * 1e-3), "unit": "pairs/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config(N, n, world),
                "e2e": {"value": pairs * e2e_steps / t_e2e, "unit": "pairs/s", "h2d_bytes_per_step": int(sh.in_bytes) * world,
                        "d2h_bytes_per_step": 8 * N * N, "ms_per_step": 1000.0 * t_e2e / e2e_steps, "steps": e2e_steps,
                        "what": "pinned host instance arrays -> H2D (every rank) -> upper-triangle row blocks -> NCCL gather -> "
                                "assemble + mirror on rank 0 -> D2H of the N x N fp64 matrix"},
                "gpu_launches": launches,
                "kernel_ms": {"discrepancy_rows_upper_rank0_share": kernel_ms, "max_over_ranks": kernel_ms_max,
                              "gather_assemble_and_gaps": ms_total / args.steps - kernel_ms_max},
                "roofline": {"kernel": "discrepancy_all_vs_all_kernel (rank 0's share of the upper triangle)", "bound": "fp64",
                             "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                             "peak_source": "measured in this run: fr3d_fp64_fma_probe (8 independent DFMA chains per thread, "
                                            "SMs x 8 blocks of 256), best of 5; MEASURED_PEAKS.json has no fp64 figure",
                             "algorithmic_flops_per_pair": flop_pair, "model": "154 n + 600 flop per instance pair, FMA = 2 "
                             "(SURVEY §8(d)); the kernel executes more (Newton on the quartic, two adjugates)"},
                "matrix_check": chk, "clocks": clocks}
        if ordering is not None:
            line["ordering"] = ordering
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
