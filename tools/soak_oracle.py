"""Soak test against the oracle: annotation sets of the CUDA path (default, single-precision screen / stacking
kernels) and of oracle/fr3d_oracle.py on many perturbed structures, the oracle in a process pool.
Usage (GPU box): python tools/soak_oracle.py <n_structures> <seed0> <sigma>"""
import multiprocessing as mp
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def oracle_one(spec):
    from conftest import ALL9
    from oracle import fr3d_oracle as O
    nts = O.build_nts(spec)
    out, c2i, tr = O.annotate(spec, ALL9, nts)
    res = {k: [(nts[a].unit_id(), nts[b].unit_id(), c) for a, b, c in v] for k, v in out.items() if v}
    return res, {k: set(v) for k, v in c2i.items() if v}, len(tr["candidates"])


def main():
    from conftest import ALL9
    from fr3d_python_b200 import synthetic
    from fr3d_python_b200.classifiers import NA_pairwise_interactions as NA
    n, seed0, sigma = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
    batch = synthetic.make_structure_batch(synthetic.pack_base(), n, seed0=seed0, sigma=sigma)
    specs = [synthetic.batch_to_spec(batch, s) for s in range(n)]
    with mp.get_context("fork").Pool(min(n, os.cpu_count() or 1)) as pool:      # before CUDA is touched
        want = pool.map(oracle_one, specs, chunksize=1)
    got, ex = NA.annotate_batch(batch, ALL9, return_raw=True)
    bad = 0
    rows = 0
    for s in range(n):
        g = ({k: list(v) for k, v in got[s][0].items() if v}, {k: set(v) for k, v in got[s][1].items() if v})
        rows += sum(len(v) for v in g[0].values())
        if g != (want[s][0], want[s][1]):
            bad += 1
            for k in set(g[0]) | set(want[s][0]):
                if g[0].get(k) != want[s][0].get(k):
                    print("structure %d label %s: gpu %d rows, oracle %d rows" % (s, k, len(g[0].get(k, [])), len(want[s][0].get(k, []))))
    print("n=%d seed0=%d sigma=%.2f: %d candidate pairs, %d rows, structures differing from the oracle: %d"
          % (n, seed0, sigma, ex["n_pairs"], rows, bad))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
