#!/usr/bin/env python
"""Benchmark of the hot path: nt-pairs classified/sec on the 1000-structure synthetic batch
(BASELINE.json configs[3]), one process per GPU, weak scaling.

  python bench.py --gpus 1 --steps K --warmup W            our arm (CUDA)
  python bench.py --impl reference ...                      the reference's CPU algorithm (oracle
                                                            port, all host cores, bounded sample)

A step = one pass of the hot path (K1 frames -> K2/K3 search -> K4 classify) over the rank's batch.
`value` is timed with inputs resident in HBM; `e2e` is the same pass through the Python/C-ABI
call with pinned HOST buffers (H2D of the nt records and D2H of the hit records inside the timed
region).  One JSON line is printed by rank 0.
"""

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CATEGORIES = {'basepair': ['cWW', 'cSS', 'cHH', 'cHS', 'cHW', 'cSH', 'cSW', 'cWH', 'cWS', 'tSS', 'tHH', 'tHS', 'tHW',
                           'tSH', 'tSW', 'tWH', 'tWS', 'tWW', 'cWB', 'cBW'],
              'basepair_detail': [], 'stacking': [], 'sO': [], 'backbone': [], 'coplanar': [], 'sugar_ribose': [],
              'covalent': [], 'near': []}
RECORD_BYTES = 760          # per-nt device record (include/fr3d_b200.h)
PAIR_BYTES = 2 * RECORD_BYTES + 12 + 32   # gather model: two nt records + (i, j, rank) + one hit slot
METRIC = "nt_pairs_classified_per_sec"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------ CPU legs --
def _oracle_one(args):
    """Annotate one synthetic structure with the oracle port; returns candidate-pair count."""
    import contextlib
    import io
    seed, n_struct_unused = args
    from fr3d_python_b200 import synthetic
    from oracle import fr3d_oracle as O
    base = synthetic.pack_base()
    tb = synthetic.make_structure_batch(base, 1, seed0=seed)
    spec = synthetic.batch_to_spec(tb, 0)
    with contextlib.redirect_stdout(io.StringIO()):
        out, c2i, tr = O.annotate(spec, CATEGORIES)
    return len(tr["candidates"])


def cpu_sample_run(pool, seeds):
    t0 = time.perf_counter()
    counts = pool.map(_oracle_one, [(s, 0) for s in seeds], chunksize=1)
    dt = time.perf_counter() - t0
    return sum(counts), dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = min(cores, 64)
    n_sample = procs               # one structure per worker per step
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        for w in range(args.warmup):
            cpu_sample_run(pool, [4000 + k for k in range(n_sample)])
        t_total, pairs_total = 0.0, 0
        for k in range(args.steps):
            pairs, dt = cpu_sample_run(pool, [4000 + k * n_sample + q for q in range(n_sample)])
            pairs_total += pairs
            t_total += dt
    value = pairs_total / t_total
    sample = "%d structures (316 nt each) of the 1000-structure batch per step, one per worker" % n_sample
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t_total / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.structures),
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "structures_per_s": n_sample * args.steps / t_total, "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, structures):
    return {"workload": "C4: synthetic batch of perturbed 1GID-scale structures (316 nt, seed 4000+k, sigma 0.10 A), "
                        "all nine annotation categories",
            "structures_per_gpu": structures, "nt_per_structure": 316,
            "l2_policy": "inputs larger than L2 (240 MB of nt records per pass vs 126 MB L2)",
            "parallelism": "structures sharded by rank, no data-path collective"}


# ------------------------------------------------------------------ clocks --
class ClockSampler(object):
    Q = "clocks.sm,clocks.max.sm,utilization.gpu,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
        "clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                clk, mx, util = float(f[0]), float(f[1]), float(f[2])
            except ValueError:
                continue
            smax = mx
            if util > 0:
                sm.append(clk)
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples_under_load": len(sm)}


# ------------------------------------------------------------------ our arm --
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--structures", type=int, default=1000, help="structures per GPU (weak scaling)")
    ap.add_argument("--cpu-structures", type=int, default=0, help="CPU baseline sample (0 = 2 per core, max 128)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chunks", type=int, default=8, help="chunks of the e2e pipeline (2 / 4 / 8 / 12 / 16: 4.06 / 3.82 / 3.71 / 3.72 / 3.88 ms)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # CPU baseline first (fork pool before any CUDA context exists): rank 0, N = 1 only
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        procs = min(cores, 64)
        n_sample = args.cpu_structures or min(2 * procs, 128)
        ctx = mp.get_context("fork")
        with ctx.Pool(procs) as pool:
            pairs, dt = cpu_sample_run(pool, [4000 + k for k in range(n_sample)])
        cpu_baseline = {"value": pairs / dt, "unit": "pairs/s", "cores": procs, "kind": "port",
                        "sample": "first %d structures of the batch (%d candidate pairs) through oracle/fr3d_oracle.py "
                                  "with a %d-process pool, %.1f s" % (n_sample, pairs, procs, dt),
                        "structures_per_s": n_sample / dt}

    import torch
    import torch.distributed as dist
    from fr3d_python_b200 import _lib, synthetic
    from fr3d_python_b200.classifiers import NA_pairwise_interactions as NA
    from fr3d_python_b200.engine import DeviceBatch, Engine, HIT_DTYPE, pin_batch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation; the contract is ONE JSON
        # line on stdout, so stdout is pointed at stderr while the communicator is built
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    sampler = ClockSampler(local_rank) if rank == 0 else None

    eng = Engine.get(local_rank)
    base = synthetic.pack_base()
    S = args.structures
    batch = synthetic.make_structure_batch(base, S, seed0=4000 + rank * S)
    focused = NA._default_focused(CATEGORIES['basepair'])
    bt, box_atoms = NA._box_table(focused, NA.load_ideal_basepair_hydrogen_bonds())
    mask = NA.category_mask(CATEGORIES)
    pinned = pin_batch(torch, batch)

    # calibration pass through the public call (also sizes the plan exactly)
    hits, n_pairs, extras = eng.annotate_batch(batch, bt, mask, pinned=pinned)
    n_hits = len(hits)
    plan = eng.make_plan(batch.n, pair_cap=n_pairs + 4096, hit_cap=n_hits + 4096)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- resident: inputs already in HBM
    db = DeviceBatch(torch, eng.device, batch, pinned)
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        eng.run_plan(db, plan, bt, mask, fit_frames=True)
    torch.cuda.synchronize()
    launches0 = eng.launches
    evs = [{k: torch.cuda.Event(enable_timing=True) for k in ("search0", "classify0", "classify1")}
           for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for k in range(args.steps):
        eng.run_plan(db, plan, bt, mask, fit_frames=True, events=evs[k])
    e1.record()
    barrier()
    gpu_launches = eng.launches - launches0
    ms_total = e0.elapsed_time(e1)
    np_chk, nh_chk = eng.read_counts(plan)
    wl = plan.counters.cpu().tolist()
    assert np_chk == n_pairs and nh_chk == n_hits, "resident pass disagrees with the calibration pass"
    classify_ms = float(np.mean([ev["classify0"].elapsed_time(ev["classify1"]) for ev in evs]))
    search_ms = float(np.mean([ev["search0"].elapsed_time(ev["classify0"]) for ev in evs]))

    # ---------------- e2e: pinned host buffers in, hit records out (per step)
    # Engine's chunked pipeline: H2D of chunk k+1 overlaps the kernels of chunk k and the D2H of k-1
    from fr3d_python_b200.engine import Pipeline
    pipe = Pipeline(eng, batch, pinned, bt, mask, n_chunks=args.chunks, pair_cap=n_pairs, hit_cap=n_hits)
    for _ in range(2):
        res_p = pipe.run()
    hits_p, n_pairs_p = res_p.hits, res_p.n_pairs
    assert n_pairs_p == n_pairs and len(hits_p) == n_hits, "pipelined pass disagrees with the calibration pass"
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        res_p = pipe.run()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    barrier()

    # host tail on a sample: hit records -> reference-format tuples (informational)
    t0 = time.perf_counter()
    n_tail = min(S, 50)
    sub = synthetic.packing.select_structures(batch, range(n_tail))
    NA.annotate_batch(sub, CATEGORIES)
    t_tail = time.perf_counter() - t0
    # ... and the same public call on a larger slice with the tail in forked worker processes (rank 0, N = 1 only:
    # the ranks of one box would share the host cores)
    t_tail_mp, n_tail_mp, tail_cores = None, min(S, 600), os.cpu_count() or 1
    if world == 1 and tail_cores > 1 and n_tail_mp >= 64:
        sub_mp = synthetic.packing.select_structures(batch, range(n_tail_mp))
        t0 = time.perf_counter()
        NA.annotate_batch(sub_mp, CATEGORIES, tail_workers=0)
        t_tail_mp = time.perf_counter() - t0

    clocks = sampler.stop() if sampler is not None else None

    # ---------------- reduce over ranks
    vals = torch.tensor([ms_total, t_e2e, float(n_pairs), float(n_hits)], dtype=torch.float64, device=eng.device)
    if world > 1:
        mx = vals.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, t_e2e = float(mx[0]), float(mx[1])
        pairs_all, hits_all = float(sm[2]), float(sm[3])
    else:
        pairs_all, hits_all = float(n_pairs), float(n_hits)
    if rank == 0:
        ms_per_step = ms_total / args.steps
        value = pairs_all * args.steps / (ms_total / 1000.0)
        e2e_value = pairs_all * args.steps / t_e2e
        peak, peak_src = load_peaks()
        achieved = n_pairs * PAIR_BYTES / (classify_ms / 1000.0) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "classify_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        h2d = pipe.h2d_bytes
        d2h = n_hits * HIT_DTYPE.itemsize + 32
        line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, S),
                "dtype_note": "every label / number of a hit record is decided in f64; f32 only screens (supersets) and "
                              "decides stacking comparisons that are certain by a margin, f64 recomputation otherwise",
                "structures_per_s": S * world * args.steps / (ms_total / 1000.0),
                "pairs_per_step": pairs_all, "hits_per_step": hits_all,
                "worklists_rank0": {"light_detail_pairs": int(wl[4]), "stacking_pairs": int(wl[5]),
                                    "pairing_pairs": int(wl[6]), "base_phosphate_pairs": int(wl[7])},
                "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "structures_per_s": S * world * args.steps / t_e2e, "ms_per_step": 1000.0 * t_e2e / args.steps,
                        "what": "engine.Pipeline.run: pinned host nt records -> H2D -> K1..K4 -> D2H hit records, "
                                "%d chunks on 3 streams" % args.chunks},
                "host_tail": {"structures_per_s_one_process": n_tail / t_tail,
                              "what": "annotate_batch incl. packing-free host post-processing to reference tuples, "
                                      "%d structures, single Python process" % n_tail,
                              "structures_per_s_all_cores": (n_tail_mp / t_tail_mp) if t_tail_mp else None,
                              "all_cores": "%d structures, tail_workers=0 (%d forked workers)" % (n_tail_mp, min(tail_cores, n_tail_mp // 8))},
                "gpu_launches": gpu_launches,
                "kernel_ms": {"classify_pairs": classify_ms, "pair_search_5_kernels": search_ms,
                              "frames_fit_and_launch_gaps": ms_per_step - classify_ms - search_ms},
                "roofline": {"kernel": "fr3d_classify_pairs = screen_records + classify_screen32 + 2 x classify_pairs_tpp + "
                                       "classify_stack32 (+ its fp64 fallback launch) + classify_list2 (eight launches: "
                                       "screen, then one or two per work list)", "bound": "hbm", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_pair": PAIR_BYTES, "model": "gather: 2 x 760 B nt records + 12 B "
                             "pair + 32 B hit slot per candidate pair"},
                "clocks": clocks}
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
