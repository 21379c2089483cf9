#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json metric: nt-pairs classified/sec & structures/sec vs host-CPU reference).

  python bench.py --gpus N --steps K --warmup W                  our arm (CUDA), workload C4
  python bench.py --workload c5 ...                              all-vs-all discrepancy of 20 000 motif instances
  python bench.py --impl reference ...                           the UNMODIFIED reference (baseline/_ref/fr3d, a copy
                                                                 of /root/reference/fr3d made by build()) on all host
                                                                 cores, bounded sample; the oracle port if absent

C4 = BASELINE configs[3]: a 1000-structure synthetic batch (perturbed 1GID-scale structures), all nine categories.
A step = one pass of the whole annotator over the batch: K1 frames -> K2/K3 search -> K4 classify -> device tail
(same-edge cleanup, emission, crossing numbers, covalent links) -> final rows in the reference's order.

  value   inputs resident in HBM, rows left in HBM (N > 1, strong scaling: gathered to rank 0's HBM with NCCL)
  e2e     packed HOST records in (pinned) -> H2D -> K1..K4 + tail -> final rows on the HOST of rank 0 (N > 1: ONE batch
          sharded by LPT over the ranks, rows gathered with a tensor collective inside the timed region)
  N = 1: scaling "weak" is moot; N > 1: "strong" by default (the one batch of configs[3] sharded over N GPUs), and the
  weak-scaling numbers (every rank its own 1000-structure batch, no exchange) are reported under "weak".

One JSON line is printed by rank 0.
"""

import argparse
import contextlib
import io
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
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def load_peaks(key="hbm_gbs", fallback=6650.0, what="B200_PROFILING.md 6.65 TB/s"):
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        if key in d:
            return float(d[key]), "measured (MEASURED_PEAKS.json %s)" % key
    return fallback, "fallback (%s)" % what


# ------------------------------------------------------------------ CPU legs --
def reference_available():
    return os.path.isdir(os.path.join(REF_DIR, "fr3d", "classifiers"))


SINGLE = {"c1": ("C1: S_1GID_full, one structure of 316 nt (the 1GID bases of the reference's Matlab golden file + template "
                 "backbone)", 1, None, None),
          "c2": ("C2: S_1S72_like, one structure of 2 844 nt (9 copies of S_1GID_full on a 3 x 3 x 1 lattice, random rotations, "
                 "sigma 0.10 A, seed 1072; no 1S72.cif exists offline)", 9, (3, 3, 1), 1072),
          "c3": ("C3: one assembly of 25 280 nt (80 copies of S_1GID_full on a 5 x 4 x 4 lattice, sigma 0.10 A, seed 25000)",
                 80, (5, 4, 4), 25000)}
_SINGLE_WORKLOAD = [None]          # set by main(): the CPU arm then times (a sample of) that one structure


def single_structure(workload):
    from fr3d_python_b200 import synthetic
    base = synthetic.pack_base()
    name, copies, lattice, seed = SINGLE[workload]
    if copies == 1:
        return base
    return synthetic.make_tiled_structure(base, copies, lattice, seed=seed, pdb=workload.upper())


def _spec_of_seed(seed):
    from fr3d_python_b200 import synthetic
    if _SINGLE_WORKLOAD[0] is not None:
        # the CPU sample of a single-structure workload: its first 316 nt (one tile copy), whatever the seed
        spec = synthetic.batch_to_spec(single_structure(_SINGLE_WORKLOAD[0]), 0)
        spec["nts"] = spec["nts"][:316]
        return spec
    base = synthetic.pack_base()
    tb = synthetic.make_structure_batch(base, 1, seed0=seed)
    return synthetic.batch_to_spec(tb, 0)


def _reference_one(seed):
    """One structure of the batch through the UNMODIFIED reference: Component construction (frame fit + hydrogens,
    fr3d/data/components.py) and annotate_nt_nt_in_structure (all nine categories).  Returns its candidate pairs."""
    sys.dont_write_bytecode = True
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    spec = _spec_of_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        from fr3d.classifiers import NA_pairwise_interactions as RNA
        from fr3d.data import Atom, Component, Structure
        comps = []
        for nt in spec["nts"]:
            atoms = [Atom(pdb=spec["pdb"], model=nt["model"], chain=nt["chain"], component_id=nt["sequence"],
                          component_number=nt["number"], component_index=nt["index"], x=a[1], y=a[2], z=a[3],
                          name=a[0], symmetry=nt["symmetry"], polymeric=True, type=a[0][0]) for a in nt["atoms"]]
            comps.append(Component(atoms, pdb=spec["pdb"], model=nt["model"], type="RNA linking", chain=nt["chain"],
                                   symmetry=nt["symmetry"], sequence=nt["sequence"], number=nt["number"],
                                   index=nt["index"], polymeric=True))
        st = Structure(comps, pdb=spec["pdb"])
        RNA.annotate_nt_nt_in_structure(st, CATEGORIES)
    c = np.array([x.centers["base"] for x in comps])
    d = np.linalg.norm(c[:, None] - c[None], axis=2)
    return int(np.triu((d >= 2.0) & (d <= 12.0), 1).sum())


def _oracle_one(seed):
    """Same through the oracle port (fallback when baseline/_ref is absent)."""
    from oracle import fr3d_oracle as O
    spec = _spec_of_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        out, c2i, tr = O.annotate(spec, CATEGORIES)
    return len(tr["candidates"])


def cpu_sample_run(pool, seeds):
    fn = _reference_one if reference_available() else _oracle_one
    t0 = time.perf_counter()
    counts = pool.map(fn, seeds, chunksize=1)
    dt = time.perf_counter() - t0
    return sum(counts), dt


def cpu_kind():
    return "reference" if reference_available() else "port"


def cpu_what():
    if reference_available():
        return "unmodified fr3d-python (baseline/_ref/fr3d): Component construction + annotate_nt_nt_in_structure"
    return "oracle/fr3d_oracle.py (port; baseline/_ref absent)"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "c5":
        import bench_c5
        return bench_c5.run_reference_arm(args)
    cores = os.cpu_count() or 1
    procs = min(cores, 64)
    if args.workload in SINGLE:    # one structure = one reference process (annotate_nt_nt_in_structure is serial)
        _SINGLE_WORKLOAD[0] = args.workload
        procs = 1
    n_sample = procs               # one structure per worker per step
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        for w in range(args.warmup):
            cpu_sample_run(pool, [4000 + k for k in range(n_sample)])
        t_total, pairs_total = 0.0, 0
        for k in range(args.steps):
            pairs, dt = cpu_sample_run(pool, [4000 + (k * n_sample + q) % args.structures for q in range(n_sample)])
            pairs_total += pairs
            t_total += dt
    value = pairs_total / t_total
    sample = "%d structures (316 nt each) of the 1000-structure batch per step, one per worker; %s" % (n_sample, cpu_what())
    if args.workload in SINGLE:
        sample = "the first 316 nt (one tile copy) of the structure per step, one process; %s" % cpu_what()
    scaling = args.scaling or ("strong" if args.gpus > 1 else "weak")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t_total / max(args.steps, 1),
            "higher_is_better": True, "scaling": scaling if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args, args.structures, args.gpus, scaling),
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": procs, "kind": cpu_kind(), "sample": sample},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "structures_per_s": n_sample * args.steps / t_total, "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, structures, world, scaling=None):
    weak = scaling == "weak" and world > 1
    heavy = getattr(args, "heavy", False)
    if args.workload in SINGLE:
        return {"workload": SINGLE[args.workload][0] + ", all nine annotation categories, final rows in the reference's order "
                            "(device tail included); a step = the whole structure",
                "structures": 1, "nt_per_structure": {"c1": 316, "c2": 2844, "c3": 25280}[args.workload],
                "l2_policy": "a 256 MB buffer is written between timed steps (the structure's records fit in L2)",
                "parallelism": "one structure runs on one GPU (SURVEY \u00a78(e)); N > 1 is not defined for this workload"}
    return {"workload": ("C4 heavy variant: synthetic batch of perturbed S_1S72_like structures (9 tiled 1GID copies, 2 844 nt, "
                         "seed 4000+k, sigma 0.10 A)" if heavy else
                         "C4: synthetic batch of perturbed 1GID-scale structures (316 nt, seed 4000+k, sigma 0.10 A)") +
                        ", all nine annotation categories, final rows in the reference's order (device tail included)",
            "structures": structures * world if weak else structures,
            "structures_per_gpu": structures if weak else structures / max(world, 1),
            "nt_per_structure": 2844 if heavy else 316,
            "l2_policy": "inputs larger than L2 at N = 1 (240 MB of nt records per pass vs 126 MB L2); a per-GPU shard "
                         "smaller than 200 MB (strong scaling at N > 1) gets a 256 MB buffer written between timed steps",
            "parallelism": "structures sharded by rank (LPT on nt count), no data-path collective; rows gathered to rank 0"}


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


def bind_to_gpu_numa_node(local_rank):
    """Pin this process (and the pinned host memory it allocates from now on) to the CPUs next to its GPU, so that the
    ranks of one box do not all stage through one NUMA node.  Best effort; returns what it did."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_words = ((os.cpu_count() or 64) + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if word >> b & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "cpus %d-%d (%d)" % (min(cpus), max(cpus), len(cpus))
    except Exception as ex:          # no NVML / no permission: stay unbound
        return "unbound (%s)" % type(ex).__name__
    return "unbound"


def init_distributed(torch, dist, local_rank):
    """NCCL prints its version banner on stdout at communicator creation; the contract is ONE JSON line on stdout,
    so stdout is pointed at stderr while the communicator is built."""
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


# ------------------------------------------------------------------ our arm --
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c4", "c5", "c1", "c2", "c3"],
                    help="c4 (default, the headline batch), c5 (discrepancy matrix), c1 / c2 / c3: ONE structure of 316 / 2 844 / "
                         "25 280 nt (BASELINE configs[0..2]; single GPU)")
    ap.add_argument("--scaling", default=None, choices=["strong", "weak"],
                    help="N > 1: strong (default) = ONE batch of --structures sharded over the ranks; weak = own batch per rank")
    ap.add_argument("--structures", type=int, default=1000, help="structures of the batch (per GPU under weak scaling)")
    ap.add_argument("--instances", type=int, default=20000, help="c5: motif instances")
    ap.add_argument("--cpu-structures", type=int, default=0, help="CPU baseline sample (0 = 2 per core, max 128)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--heavy", action="store_true",
                    help="C4 heavy variant (SURVEY §8(d)): every structure is S_1S72_like, 9 tiled copies = 2 844 nt (implies "
                         "--no-cpu-baseline; not the headline configuration)")
    ap.add_argument("--chunks", type=int, default=10, help="chunks of the e2e pipeline per GPU")
    ap.add_argument("--taper", default="0.3,0.7,1.2,1.3,1.3,1.3,1.3,1.2,0.9,0.5",
                    help="relative chunk sizes of the e2e pipeline (comma separated, as many as --chunks; 'none' = equal chunks)")
    ap.add_argument("--streams", type=int, default=3, help="streams the chunks of the e2e pipeline rotate over")
    ap.add_argument("--no-graph", action="store_true", help="time Pipeline.run (one Python launch per kernel) instead of the CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if args.heavy:
        args.no_cpu_baseline = True
    if args.workload in SINGLE:
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            if int(os.environ.get("RANK", "0")) == 0:
                print(json.dumps({"metric": METRIC, "workload": args.workload, "unavailable": "a single structure runs on one GPU"}))
            return
        _SINGLE_WORKLOAD[0] = args.workload
        args.structures, args.chunks, args.cpu_structures = 1, 1, 1
    if args.workload == "c5":
        import bench_c5
        return bench_c5.main(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    scaling = args.scaling or ("strong" if world > 1 else "weak")

    # CPU baseline first (fork pool before any CUDA context exists): rank 0, N = 1 only
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        procs = 1 if args.workload in SINGLE else min(cores, 64)
        n_sample = args.cpu_structures or min(2 * procs, 128)
        ctx = mp.get_context("fork")
        with ctx.Pool(procs) as pool:
            pairs, dt = cpu_sample_run(pool, [4000 + k for k in range(n_sample)])
        cpu_baseline = {"value": pairs / dt, "unit": "pairs/s", "cores": procs, "kind": cpu_kind(),
                        "sample": ("the first 316 nt (one tile copy, %d candidate pairs) of the structure, %s, one process, %.1f s"
                                   % (pairs, cpu_what(), dt)) if args.workload in SINGLE else
                                  "first %d structures of the batch (%d candidate pairs), %s, %d-process pool, %.1f s"
                                  % (n_sample, pairs, cpu_what(), procs, dt),
                        "structures_per_s": n_sample / dt}

    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "single rank: unbound"
    import torch
    import torch.distributed as dist
    from fr3d_python_b200 import packing, parallel, synthetic
    from fr3d_python_b200 import tail as tailmod
    from fr3d_python_b200.classifiers import NA_pairwise_interactions as NA
    from fr3d_python_b200.engine import DeviceBatch, DeviceIdent, Engine, Pipeline, pin_batch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        init_distributed(torch, dist, local_rank)
    sampler = ClockSampler(local_rank) if rank == 0 else None

    eng = Engine.get(local_rank)
    dev = eng.device
    base = synthetic.pack_base()
    if args.heavy:
        base = synthetic.make_tiled_structure(base, 9, (3, 3, 1), seed=1072, pdb="S1S72")
    S = args.structures
    focused = NA._default_focused(CATEGORIES['basepair'])
    bt, box_atoms, labels = NA._tables(focused, NA.load_ideal_basepair_hydrogen_bonds())
    mask = NA.category_mask(CATEGORIES)
    l2_buf = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)         # 256 MB > 126 MB L2
    shared = {"full_batch": None, "rows_expected": None}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(batch_local, mine, n_structures_total, gather):
        """Resident and end-to-end timing of one sharding of the work.  batch_local: this rank's structures;
        gather: rows go to rank 0 inside the timed regions (strong scaling).  Returns a dict of local measurements."""
        out = {}
        pinned = pin_batch(torch, batch_local)
        ident = DeviceIdent(torch, dev, tailmod.tail_ident(batch_local))
        tail = (ident, labels)
        # calibration pass through the public call (also sizes the plan exactly)
        hits, n_pairs, extras = eng.annotate_batch(batch_local, bt, mask, pinned=pinned, labels=labels, want_hits=True)
        n_hits, n_rows = len(hits), len(extras["rows"])
        plan = eng.make_plan(batch_local.n, pair_cap=n_pairs + 4096, hit_cap=n_hits + 4096,
                             tail=(ident.n_struct, ident.n_ends), row_cap=n_rows + 4096)
        out.update(n_pairs=n_pairs, n_hits=n_hits, n_rows=n_rows)
        mine_l = list(mine)
        db = DeviceBatch(torch, dev, batch_local, pinned)

        def resident_step(events=None):
            eng.run_plan(db, plan, bt, mask, fit_frames=True, events=events, tail=tail)

        def timed_loop(step):
            step_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            barrier()
            for k in range(args.steps):
                if small:
                    l2_buf.add_(1)
                step_ev[k][0].record()
                step(k)
                step_ev[k][1].record()
            barrier()
            return float(sum(a.elapsed_time(b) for a, b in step_ev))

        # ---------------- resident: inputs already in HBM (per-kernel events; local shard, nothing exchanged)
        torch.cuda.synchronize()
        for _ in range(args.warmup):
            resident_step()
        torch.cuda.synchronize()
        launches0 = eng.launches
        names = ("search0", "classify0", "classify1", "tail1")
        evs = [{k: torch.cuda.Event(enable_timing=True) for k in names} for _ in range(args.steps)]
        small = batch_local.nbytes_device() < 200e6                # shard not much larger than L2: flush between steps
        out["ms_total"] = timed_loop(lambda k: resident_step(events=evs[k]))
        out["gpu_launches"] = eng.launches - launches0
        np_chk, nh_chk = eng.read_counts(plan)
        nr_chk = eng.read_tail_counts(plan)
        assert (np_chk, nh_chk, nr_chk) == (n_pairs, n_hits, n_rows), \
            "resident pass (pairs, hits, rows) = %r disagrees with the calibration pass %r" % (
                (np_chk, nh_chk, nr_chk), (n_pairs, n_hits, n_rows))
        out["worklists"] = plan.counters.cpu().tolist()
        out["classify_ms"] = float(np.mean([ev["classify0"].elapsed_time(ev["classify1"]) for ev in evs]))
        out["search_ms"] = float(np.mean([ev["search0"].elapsed_time(ev["classify0"]) for ev in evs]))
        out["tail_ms"] = float(np.mean([ev["classify1"].elapsed_time(ev["tail1"]) for ev in evs]))
        out["l2_flush_between_steps"] = bool(small)

        # ---------------- e2e: pinned host buffers in, final rows on the host of rank 0 out (per step)
        if gather:
            chunks = max(1, args.chunks // world)
            sh = parallel.ShardedAnnotator(shared["full_batch"], CATEGORIES, rank=rank, world_size=world,
                                           device=local_rank, n_chunks=chunks)
            for _ in range(2):
                res = sh.run()
            if rank == 0:
                assert res.n_rows == shared["rows_expected"], "sharded pass disagrees with the single-GPU row count"
            # `value` under strong scaling: shard kernels on the resident records + NCCL gather into rank 0's HBM
            # (one chunk per rank: nothing to overlap when the records are resident)
            sh1 = parallel.ShardedAnnotator(shared["full_batch"], CATEGORIES, rank=rank, world_size=world,
                                            device=local_rank, n_chunks=1)
            for _ in range(2):
                res1 = sh1.run()
            if rank == 0:
                assert res1.n_rows == shared["rows_expected"]
            launches0 = eng.launches
            out["ms_total"] = timed_loop(lambda k: sh1.enqueue(upload=False))
            out["gpu_launches"] = eng.launches - launches0
            del sh1
            barrier()
            t0 = time.perf_counter()
            for k in range(args.steps):
                res = sh.run()
            torch.cuda.synchronize()
            out["t_e2e"] = time.perf_counter() - t0
            barrier()
            out["h2d"] = sh.pipe.h2d_bytes if sh.pipe is not None else 0
            out["d2h"] = (4 * sh.block_ints * world) if rank == 0 else 0
            out["e2e_what"] = "parallel.ShardedAnnotator.run: LPT shard -> per rank pinned records -> H2D -> K1..K4 + " \
                              "device tail (%d chunks, rows written straight into the rank's send block) -> one NCCL " \
                              "gather of fixed-size blocks to rank 0 -> one D2H" % sh.n_chunks
            out["sharded_result"] = res
        else:
            weights = None
            if args.taper != "none":
                weights = [float(v) for v in args.taper.split(",")]
                weights = weights if len(weights) == min(args.chunks, batch_local.n_structures) else None
            pipe = Pipeline(eng, batch_local, pinned, bt, mask, n_chunks=args.chunks, pair_cap=n_pairs, hit_cap=n_hits,
                            labels=labels, row_cap=n_rows, chunk_weights=weights, n_streams=args.streams)
            for _ in range(2):
                res = pipe.run()
            assert res.n_pairs == n_pairs and len(res.rows) == n_rows, "pipelined pass disagrees with the calibration pass"
            use_graph = not args.no_graph
            if use_graph:
                try:
                    res = pipe.run_graph()
                    assert res.n_pairs == n_pairs and res.n_rows == n_rows, "graph replay disagrees with the calibration pass"
                except Exception as ex:                     # capture not possible here: the plain pipeline is timed
                    sys.stderr.write("CUDA graph capture failed (%r); timing Pipeline.run instead\n" % (ex,))
                    use_graph = False
            step = pipe.run_graph if use_graph else pipe.run
            barrier()
            t0 = time.perf_counter()
            for k in range(args.steps):
                res = step()
            torch.cuda.synchronize()
            out["t_e2e"] = time.perf_counter() - t0
            barrier()
            out["h2d"] = pipe.h2d_bytes
            out["d2h"] = pipe.graph_d2h_bytes if use_graph else (n_rows * 16 + 8 * batch_local.n_structures + 96 * len(pipe.chunks))
            out["e2e_what"] = "engine.Pipeline.%s: pinned host nt records -> H2D -> K1..K4 + device tail -> D2H of the final rows " \
                              "(label id, nt1, nt2, crossing; reference order), %d chunks%s on %d streams%s" % (
                                  "run_graph" if use_graph else "run", args.chunks,
                                  " (relative sizes %s: small first and last chunks)" % args.taper if weights else "", args.streams,
                                  ", the whole pass captured in ONE CUDA graph and replayed" if use_graph else "")
            if use_graph:          # the plain pipeline beside it
                barrier()
                t0 = time.perf_counter()
                for k in range(args.steps):
                    pipe.run()
                torch.cuda.synchronize()
                out["t_e2e_no_graph"] = time.perf_counter() - t0
        return out

    def reduce_max_sum(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            mx, sm = t.clone(), t.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            return mx.tolist(), sm.tolist()
        return t.tolist(), t.tolist()

    results = {}
    if world == 1 or scaling == "strong":
        # ONE batch of S structures; N > 1: sharded by LPT, rows gathered to rank 0
        full_batch = single_structure(args.workload) if args.workload in SINGLE else \
            synthetic.make_structure_batch(base, S, seed0=4000)
        shared["full_batch"] = full_batch
        parts = parallel.lpt_partition(np.diff(full_batch.struct_off).astype(int).tolist(), world)
        mine = parts[rank]
        local = packing.select_structures(full_batch, mine) if world > 1 else full_batch
        ref = None
        if world > 1 and rank == 0:          # the single-GPU answer the sharded run must reproduce
            ref = NA.annotate_batch(full_batch, CATEGORIES, as_arrays=True)
            shared["rows_expected"] = ref.n_rows
        m = measure(local, mine, S, gather=(world > 1))
        got = m.pop("sharded_result", None)
        if ref is not None:
            same = all(np.array_equal(got.rows_of(s), ref.rows_of(s)) for s in range(S))
            assert same, "rank-0 result of the sharded run differs from the single-GPU run"
            m["sharded_equals_single_gpu"] = True
        results["main"] = m
    if world > 1:
        # weak scaling: every rank its own S-structure batch, nothing exchanged
        own = synthetic.make_structure_batch(base, S, seed0=4000 + rank * S)
        results["weak"] = measure(own, list(range(S)), S, gather=False)
        if scaling == "weak":
            results["main"] = results["weak"]

    # ---------------- through the public API (rank 0, N = 1): array API and tuples; packing included
    api = None
    if world == 1 and rank == 0 and not args.heavy and args.workload not in SINGLE:
        batch = shared["full_batch"]
        pinned = pin_batch(torch, batch)
        NA.annotate_batch(batch, CATEGORIES, pinned=pinned, as_arrays=True)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            res = NA.annotate_batch(batch, CATEGORIES, pinned=pinned, as_arrays=True)
        t_arr = (time.perf_counter() - t0) / reps
        n_t = min(S, 200)
        t0 = time.perf_counter()
        for s in range(n_t):
            res[s]
        t_tuples = (time.perf_counter() - t0) / n_t
        # Structure-level input: column data (what an atom_site reader yields) -> pack_columns -> pipeline
        n_in = min(S, 100)
        specs = [synthetic.batch_to_spec(batch, s) for s in range(n_in)]
        cols = packing.specs_to_columns(specs)
        t0 = time.perf_counter()
        b_in = packing.pack_columns(*cols)
        t_pack = time.perf_counter() - t0
        t0 = time.perf_counter()
        NA.annotate_batch(b_in, CATEGORIES, as_arrays=True)
        t_ann = time.perf_counter() - t0
        # CIF text in: native mmCIF ingest (C++, one thread per file) -> pinned records -> annotate_batch array API
        from fr3d_python_b200.cif import native as cif_native, writer as cif_writer
        ident_op = [{"id": "1", "name": "1_555", "matrix": [[1, 0, 0], [0, 1, 0], [0, 0, 1]], "vector": [0, 0, 0]}]
        n_distinct = min(S, 32)
        texts = [cif_writer.spec_to_cif(synthetic.batch_to_spec(batch, s), operators=ident_op, assembly=[("1", "A,B")]).encode()
                 for s in range(n_distinct)]
        texts_all = [texts[k % n_distinct] for k in range(S)]
        cif_native.cif_to_batch(texts_all[:8])
        from fr3d_python_b200.engine import pinned_planes
        t0 = time.perf_counter()
        b_cif = cif_native.cif_to_batch(texts_all, n_threads=0, pinned=True)
        t_ingest = time.perf_counter() - t0
        t0 = time.perf_counter()
        res_cif = NA.annotate_batch(b_cif, CATEGORIES, pinned=pinned_planes(b_cif), as_arrays=True)
        t_cif_ann = time.perf_counter() - t0
        api = {"annotate_batch_array_api_structures_per_s": S / t_arr,
               "e2e_cif_in": {"structures_per_s": S / (t_ingest + t_cif_ann), "ingest_ms_per_structure": 1e3 * t_ingest / S,
                              "annotate_ms_per_structure": 1e3 * t_cif_ann / S,
                              "cif_bytes_per_structure": len(texts[0]), "host_threads": os.cpu_count(), "rows": res_cif.n_rows,
                              "what": "mmCIF TEXT of %d structures (%d distinct files written from the batch, 3-decimal "
                                      "coordinates) -> cif.native (C++ tokenizer, residue delimitation like fr3d/cif/reader.py, "
                                      "slot packing; one thread per file) -> pinned records -> annotate_batch array API" % (S, n_distinct)},
               "annotate_batch_array_api_what": "NA.annotate_batch(batch, categories, pinned=..., as_arrays=True): one "
                                                "un-pipelined pass incl. buffer allocation, %d structures" % S,
               "tuples_structures_per_s_one_process": 1.0 / (t_arr / S + t_tuples),
               "tuples_what": "array API + materialising the reference's dict of unit-id tuples for every structure "
                              "(%.3f ms per structure in Python, measured on %d)" % (1e3 * t_tuples, n_t),
               "e2e_structure_in": {"structures_per_s": n_in / (t_pack + t_ann), "pack_ms_per_structure": 1e3 * t_pack / n_in,
                                    "what": "atom columns of %d structures -> packing.pack_columns (host numpy) -> "
                                            "annotate_batch array API; packing included" % n_in}}

    clocks = sampler.stop() if sampler is not None else None

    # ---------------- reduce over ranks
    def summarise(m, S_total):
        mx, sm = reduce_max_sum([m["ms_total"], m["t_e2e"], float(m["n_pairs"]), float(m["n_hits"]), float(m["n_rows"]),
                                 float(m["h2d"]), float(m["d2h"]), m["classify_ms"], m["search_ms"], m["tail_ms"],
                                 float(m["gpu_launches"])])
        ms_total, t_e2e = mx[0], mx[1]
        pairs_all, hits_all, rows_all = sm[2], sm[3], sm[4]
        return {"ms_per_step": ms_total / args.steps, "value": pairs_all * args.steps / (ms_total / 1000.0),
                "structures_per_s": S_total * args.steps / (ms_total / 1000.0),
                "e2e_value": pairs_all * args.steps / t_e2e, "e2e_structures_per_s": S_total * args.steps / t_e2e,
                "e2e_ms_per_step": 1000.0 * t_e2e / args.steps, "pairs": pairs_all, "hits": hits_all, "rows": rows_all,
                "h2d": sm[5], "d2h": sm[6], "classify_ms_max": mx[7], "search_ms_max": mx[8], "tail_ms_max": mx[9],
                "gpu_launches": int(sm[10])}

    strong = world == 1 or scaling == "strong"
    main_s = summarise(results["main"], S if strong else S * world)
    weak_s = summarise(results["weak"], S * world) if ("weak" in results and scaling == "strong") else None
    if rank == 0:
        m = results["main"]
        peak, peak_src = load_peaks()
        achieved = m["n_pairs"] * PAIR_BYTES / (m["classify_ms"] / 1000.0) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "classify_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        wl = m["worklists"]
        line = {"metric": METRIC, "value": main_s["value"], "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": main_s["ms_per_step"], "higher_is_better": True,
                "scaling": scaling if world > 1 else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, S, world, scaling),
                "dtype_note": "every label / number of a hit record is decided in f64; f32 only screens (supersets) and "
                              "decides stacking comparisons that are certain by a margin, f64 recomputation otherwise; "
                              "the tail is integer work",
                "structures_per_s": main_s["structures_per_s"],
                "pairs_per_step": main_s["pairs"], "hits_per_step": main_s["hits"], "rows_per_step": main_s["rows"],
                "worklists_rank0": {"light_detail_pairs": int(wl[4]), "stacking_pairs": int(wl[5]),
                                    "pairing_pairs": int(wl[6]), "base_phosphate_pairs": int(wl[7])},
                "e2e": {"value": main_s["e2e_value"], "unit": "pairs/s", "h2d_bytes_per_step": int(main_s["h2d"]),
                        "d2h_bytes_per_step": int(main_s["d2h"]), "structures_per_s": main_s["e2e_structures_per_s"],
                        "ms_per_step": main_s["e2e_ms_per_step"], "what": m["e2e_what"]},
                "e2e_without_cuda_graph_ms_per_step": (1000.0 * m["t_e2e_no_graph"] / args.steps) if "t_e2e_no_graph" in m else None,
                "gpu_launches": main_s["gpu_launches"],
                "kernel_ms": {"classify_pairs": m["classify_ms"], "pair_search_5_kernels": m["search_ms"],
                              "device_tail_5_kernels": m["tail_ms"],
                              "frames_fit_gather_and_launch_gaps": main_s["ms_per_step"] - m["classify_ms"] - m["search_ms"] - m["tail_ms"]},
                "roofline": {"kernel": "fr3d_classify_pairs = screen_records + classify_screen32 + 2 x classify_pairs_tpp + "
                                       "classify_stack32 (+ its fp64 fallback launch) + classify_list2 (eight launches: "
                                       "screen, then one or two per work list), rank 0", "bound": "hbm", "achieved": achieved,
                             "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                             "peak_source": peak_src, "algorithmic_bytes_per_pair": PAIR_BYTES,
                             "model": "gather: 2 x 760 B nt records + 12 B pair + 32 B hit slot per candidate pair"},
                "l2_flush_between_steps": m["l2_flush_between_steps"],
                "numa_binding_rank0": numa, "clocks": clocks}
        if m.get("sharded_equals_single_gpu"):
            line["sharded_equals_single_gpu"] = True
        if weak_s is not None:
            line["weak"] = {"value": weak_s["value"], "ms_per_step": weak_s["ms_per_step"],
                            "structures_per_s": weak_s["structures_per_s"], "structures_per_gpu": S,
                            "e2e": {"value": weak_s["e2e_value"], "ms_per_step": weak_s["e2e_ms_per_step"],
                                    "structures_per_s": weak_s["e2e_structures_per_s"],
                                    "h2d_bytes_per_step": int(weak_s["h2d"]), "d2h_bytes_per_step": int(weak_s["d2h"])},
                            "what": "weak scaling: every rank its own %d-structure batch, nothing exchanged" % S}
        if api is not None:
            line["api"] = api
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
