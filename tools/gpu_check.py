"""First-run triage on a GPU box: CUDA path vs oracle vs golden, with detailed mismatch prints.
Usage (under gpurun):  python tools/gpu_check.py > gpurun_out/check.log 2>&1
"""

import gzip
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import fr3d_oracle as O  # noqa: E402
from fr3d_python_b200 import packing, synthetic, postprocess  # noqa: E402
from fr3d_python_b200.classifiers import NA_pairwise_interactions as NA  # noqa: E402
from fr3d_python_b200.geometry import superpositions, discrepancy as disc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def load(name):
    with gzip.open(os.path.join(GOLD, name), "rt") as f:
        return json.load(f)


def hits_to_traces(hits, lo=0):
    """Per-slot dictionaries {(nt1, nt2): label} in the oracle's trace vocabulary."""
    out = {"sO": {}, "stacking": {}, "sugar_ribose": {}, "bph": {}, "br": {}, "cp": set(), "bp": {}}
    for h in hits.tolist():
        i, j, rank, slot, code, box, cd, gap = h
        a, b = i - lo, j - lo
        if slot == 0:
            out["sO"][(a, b)] = postprocess.so_labels(code)[0]
        elif slot == 1:
            out["sO"][(b, a)] = postprocess.so_labels(code)[0]
        elif slot == 2:
            out["stacking"][(a, b)] = postprocess.stack_labels(code)[0]
        elif slot == 3:
            out["sugar_ribose"][(a, b)] = ["cSR", "tSR"][code]
        elif slot == 4:
            out["sugar_ribose"][(b, a)] = ["cSR", "tSR"][code]
        elif slot == 5:
            out["bph"][(a, b)] = ("n" if code & 16 else "") + str(code & 15) + "BPh"
        elif slot == 6:
            out["br"][(a, b)] = ("n" if code & 16 else "") + str(code & 15) + "BR"
        elif slot == 7:
            out["cp"].add((a, b))
        elif slot == 8:
            out["bp"][(a, b)] = (code, box, cd, gap)
    return out


def compare_annotation(spec, categories, label, golden=None):
    print("=== %s: %d nts" % (label, len(spec["nts"])))
    t = time.time()
    oi, oc, otr = O.annotate(spec, categories)
    t_or = time.time() - t
    batch = packing.pack_specs(spec)
    t = time.time()
    results, extras = NA.annotate_batch(batch, categories, return_raw=True)
    t_gpu = time.time() - t
    gi, gc = results[0]
    print("oracle %.2fs, gpu path %.3fs, pairs %d (oracle %d), hits %d" % (
        t_or, t_gpu, extras["n_pairs"], len(otr["candidates"]), len(extras["hits"])))
    # frames
    nts = O.build_nts(spec)
    dc = max(np.abs(extras["center"][k] - nts[k].base_center).max() for k in range(len(nts)))
    dr = max(np.abs(extras["rot"][k].reshape(3, 3) - nts[k].rotation).max() for k in range(len(nts)))
    print("frames: max |dcentre| %.3e  max |dR| %.3e" % (dc, dr))
    # candidates
    pi, pj, pr = extras["pairs"]
    order = np.lexsort((pj, pi, pr))
    cand = list(zip(pi[order].tolist(), pj[order].tolist()))
    ocand = [tuple(p) for p in otr["candidates"]]
    print("candidates equal (ordered): %s  set-equal: %s" % (cand == ocand, set(cand) == set(ocand)))
    if set(cand) != set(ocand):
        print("  only gpu:", sorted(set(cand) - set(ocand))[:10])
        print("  only oracle:", sorted(set(ocand) - set(cand))[:10])
    # per-rule traces
    ht = hits_to_traces(extras["hits"])
    for key, okey in (("sO", "sO"), ("stacking", "stacking"), ("sugar_ribose", "sugar_ribose")):
        if okey not in otr or not otr[okey]:
            continue
        od = {(a, b): l for a, b, l in otr[okey] if l}
        gd = ht[key]
        bad = [(k, od.get(k), gd.get(k)) for k in set(od) | set(gd) if od.get(k) != gd.get(k)]
        print("%s: oracle %d gpu %d mismatches %d" % (key, len(od), len(gd), len(bad)))
        for b in bad[:8]:
            print("   ", b)
    if otr["backbone"]:
        od = {(a, b): l for a, b, l, r in otr["backbone"] if l}
        bad = [(k, od.get(k), ht["bph"].get(k)) for k in set(od) | set(ht["bph"]) if od.get(k) != ht["bph"].get(k)]
        print("BPh: oracle %d gpu %d mismatches %d" % (len(od), len(ht["bph"]), len(bad)))
        for b in bad[:8]:
            print("   ", b)
        od = {(a, b): r for a, b, l, r in otr["backbone"] if r}
        bad = [(k, od.get(k), ht["br"].get(k)) for k in set(od) | set(ht["br"]) if od.get(k) != ht["br"].get(k)]
        print("BR: oracle %d gpu %d mismatches %d" % (len(od), len(ht["br"]), len(bad)))
        for b in bad[:8]:
            print("   ", b)
    # final
    og = {k: [(nts[a].unit_id(), nts[b].unit_id(), c) for a, b, c in v] for k, v in oi.items()}
    gg = {k: [tuple(t) for t in v] for k, v in gi.items()}
    same_order = og == gg
    same_set = {k: sorted(v, key=str) for k, v in og.items() if v} == {k: sorted(v, key=str) for k, v in gg.items() if v}
    print("FINAL ordered-equal: %s   set-equal: %s   rows %d" % (same_order, same_set, sum(len(v) for v in gg.values())))
    if not same_set:
        for k in sorted(set(og) | set(gg)):
            a, b = set(og.get(k, [])), set(gg.get(k, []))
            if a != b:
                print("   %-8s only oracle %s | only gpu %s" % (k, sorted(a - b, key=str)[:4], sorted(b - a, key=str)[:4]))
    print("category_to_interactions equal:", {k: set(v) for k, v in oc.items() if v} == {k: set(v) for k, v in gc.items() if v})
    if golden is not None:
        gold = load(golden)
        uid = gold["unit_ids"]
        want = {k: [(uid[a], uid[b], c) for a, b, c in v] for k, v in gold["interaction_to_list_of_tuples"].items()}
        print("GOLDEN (reference) ordered-equal:", want == gg)
    return same_set


def check_geometry():
    geo = load("geometry.json.gz")
    worst = 0.0
    for k in geo["kabsch"]:
        U, new1, mean1, rmsd, sse, mean2 = superpositions.besttransformation(np.array(k["set1"]), np.array(k["set2"]))
        dU = np.abs(np.asarray(U) - np.array(k["U"])).max()
        dn = np.abs(np.asarray(new1) - np.array(k["new1"])).max()
        ds = abs(sse - k["sse"])
        Uw, n1w, m1w, rmsdw, ssew = superpositions.besttransformation_weighted(np.array(k["set1"]), np.array(k["set2"]),
                                                                              k["weights"])
        dUw = np.abs(np.asarray(Uw) - np.array(k["Uw"])).max()
        worst = max(worst, dU, dn, dUw)
        if max(dU, dn, dUw) > 1e-9 or ds > 1e-9:
            print("  kabsch %s dU %.2e dnew1 %.2e dsse %.2e dUw %.2e" % (k["name"], dU, dn, ds, dUw))
    print("kabsch KATs: %d problems, worst |d| %.3e" % (len(geo["kabsch"]), worst))
    for k in geo["discrepancy_kats"]:
        c1 = [np.array(c) for c in k["centers1"]]
        r1 = [np.array(c) for c in k["rotations1"]]
        c2 = [np.array(c) for c in k["centers2"]]
        r2 = [np.array(c) for c in k["rotations2"]]
        d = disc.matrix_discrepancy(c1, r1, c2, r2)
        print("  discrepancy %-8s gpu %.16f ref %.16f  diff %.2e" % (k["name"], d, k["d"], abs(d - k["d"])))
        if "cutoff_2.5" in k:
            hi = disc.matrix_discrepancy_cutoff(c1, r1, c2, r2, 2.5)
            lo = disc.matrix_discrepancy_cutoff(c1, r1, c2, r2, 0.5)
            print("     cutoff 2.5 -> %s (ref %s), 0.5 -> %s (ref %s)" % (hi, k["cutoff_2.5"], lo, k["cutoff_0.5"]))
    av = geo["all_vs_all"]
    C, R = synthetic.make_instances(av["base_centers"], av["base_rotations"], av["N"], av["seed"])
    print("instances reproduce golden inputs:", np.array_equal(C, np.array(av["centers"])),
          np.array_equal(R, np.array(av["rotations"])))
    D = disc.all_vs_all(C, R)
    Dref = np.array(av["D"])
    print("all-vs-all N=%d: max |dD| %.3e  symmetric %s  diag0 %s" % (av["N"], np.abs(D - Dref).max(),
                                                                   np.array_equal(D, D.T), np.all(np.diag(D) == 0)))
    has = np.ones((av["N"], av["n"]), dtype=np.uint8)
    has[:, 0] = 0
    Dn = disc.all_vs_all(C, R, has_rot=has)
    print("all-vs-all no-rot0: max |dD| %.3e" % np.abs(Dn - np.array(av["D_norot0"])).max())
    row = disc.one_vs_many(C[0], R[0], C, R, cutoff=av["cutoff"])
    ref = np.array([np.nan if v is None else v for v in av["row0_cutoff"]])
    print("one-vs-many cutoff: None pattern equal %s, max |d| %.3e" % (
        np.array_equal(np.isnan(row), np.isnan(ref)), np.nanmax(np.abs(row - ref))))


def main():
    all9 = load("annot_1gid_full_all9.json.gz")["categories"]
    ok = True
    ok &= compare_annotation(load("spec_1s72_strand.json.gz"), all9, "1S72 strand", "annot_1s72_strand_all9.json.gz")
    ok &= compare_annotation(load("spec_1gid_full.json.gz"), all9, "1GID full all9", "annot_1gid_full_all9.json.gz")
    ok &= compare_annotation(load("spec_1gid_base.json.gz"), {'basepair': NA.Leontis_Westhof_basepairs},
                             "1GID base bp only", "annot_1gid_base_bp.json.gz")
    base = synthetic.pack_base()
    tb = synthetic.make_structure_batch(base, 3, seed0=4000)
    for s in range(3):
        ok &= compare_annotation(synthetic.batch_to_spec(tb, s), all9, "perturbed copy %d" % s)
    check_geometry()
    print("ALL ANNOTATION SETS EQUAL:", ok)


if __name__ == "__main__":
    main()
