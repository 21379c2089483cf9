"""Generate tests/golden/adversarial.json.gz (DEV CONTAINER ONLY: imports the unmodified reference from
/root/reference to record ITS answer at every probe point).

    python tools/make_adversarial.py

1. pairs of 1GID-full are chosen per rule family from the oracle's traces;
2. every (pair, motion) is scanned on a grid, every change of the oracle's annotation (rows + number of candidate
   pairs) between two grid points is bisected to |hi - lo| <= 1e-13, and the boundary is attributed to the threshold
   whose margin changes sign there (oracle.PROBE);
3. boundaries are kept so that every threshold name that occurs gets up to PER_THRESHOLD of them;
4. at lo - d and hi + d, d in tests/adversarial.DELTAS, the oracle and the unmodified reference annotate the pair; the
   reference's rows go into the golden file together with t and a coordinate checksum;
5. "cell" cases: pairs shifted as a whole so that a base centre coordinate sits at a cube boundary 24 -+ d (NA:424-436).
"""

import contextlib
import gzip
import io
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import adversarial as A  # noqa: E402
import ref_fixtures  # noqa: E402
from conftest import ALL9, load_golden  # noqa: E402
from oracle import fr3d_oracle as O  # noqa: E402

PER_THRESHOLD = 5
MAX_PER_SCAN = 5


def reference_rows(spec):
    """The unmodified reference on a spec: {label: [[pos1, pos2, crossing], ...]} without covalent links."""
    ref_fixtures.add_reference_to_path()
    from fr3d.classifiers import NA_pairwise_interactions as RNA
    with contextlib.redirect_stdout(io.StringIO()):
        st = ref_fixtures.ref_structure_from_spec(spec)
        i2t, c2i, _, _ = RNA.annotate_nt_nt_in_structure(st, ALL9)
    uid = {c.unit_id(): k for k, c in enumerate(st.residues(type=["RNA linking", "DNA linking"]))}
    return {k: [[uid[a], uid[b], c] for a, b, c in v] for k, v in i2t.items() if v and not k.startswith("p_")}


def evaluate(case, motion, t, probe=False, shift=None):
    spec = case.moved(motion, t, shift)
    O.PROBE = [] if probe else None
    try:
        out, c2i, tr = O.annotate(spec, ALL9)
        margins = O.PROBE
    finally:
        O.PROBE = None
    rows = {k: [list(x) for x in v] for k, v in out.items() if v and not k.startswith("p_")}
    sig = (A.signature(rows), len(tr["candidates"]))
    return spec, rows, len(tr["candidates"]), sig, margins


def choose_pairs(spec):
    out, c2i, tr = O.annotate(spec, ALL9)
    chosen = {}

    def take(family, pairs, n):
        k = 0
        for i, j in pairs:
            key = (min(i, j), max(i, j))
            if key not in chosen:
                chosen[key] = family
                k += 1
                if k >= n:
                    break

    take("sO", [(i, j) for i, j, a in tr["sO"] if a and not a.startswith("n")], 3)
    take("near sO", [(i, j) for i, j, a in tr["sO"] if a.startswith("n")], 2)
    take("stacking", [(i, j) for i, j, a in tr["stacking"] if a and not a.startswith("n")][::40], 5)
    take("near stacking", [(i, j) for i, j, a in tr["stacking"] if a.startswith("n")][::6], 3)
    take("sugar ribose", [(i, j) for i, j, a in tr["sugar_ribose"] if a], 3)
    take("base phosphate", [(i, j) for i, j, bph, br in tr["backbone"] if bph and not bph.startswith("n")], 3)
    take("near base phosphate", [(i, j) for i, j, bph, br in tr["backbone"] if bph.startswith("n")], 2)
    take("base ribose", [(i, j) for i, j, bph, br in tr["backbone"] if br], 3)
    seen_lw = set()
    bp = []
    for i, j, lw, sub, cd, gap in tr["basepair"]:
        if lw and lw not in seen_lw:
            seen_lw.add(lw)
            bp.append((i, j))
    take("basepair", bp, 14)
    take("coplanar", [(i, j) for i, j, cp, g, md, v in tr["coplanar"] if cp][::25], 3)
    quiet = set((min(i, j), max(i, j)) for i, j in tr["candidates"])
    busy = set()
    for k, v in out.items():
        if not k.startswith("p_"):
            busy.update((min(a, b), max(a, b)) for a, b, c in v)
    take("no interaction", sorted(quiet - busy)[::300], 4)
    return chosen


def main():
    t_start = time.time()
    spec = load_golden("spec_1gid_full.json.gz")
    chosen = choose_pairs(spec)
    print("pairs:", len(chosen), sorted(set(chosen.values())))
    boundaries = []
    for (i, j), family in sorted(chosen.items()):
        case = A.PairCase(spec, i, j)
        for motion, (t0, t1, npts, unit) in A.MOTIONS.items():
            ts = np.linspace(t0, t1, npts)
            sigs = [evaluate(case, motion, float(t))[3] for t in ts]
            changes = [k for k in range(npts - 1) if sigs[k] != sigs[k + 1]]
            if len(changes) > MAX_PER_SCAN:
                changes = [changes[int(round(q * (len(changes) - 1) / (MAX_PER_SCAN - 1)))] for q in range(MAX_PER_SCAN)]
            for k in sorted(set(changes)):
                lo, hi = float(ts[k]), float(ts[k + 1])
                s_lo = sigs[k]
                for _ in range(80):
                    if hi - lo <= 1e-13:
                        break
                    mid = 0.5 * (lo + hi)
                    if mid <= lo or mid >= hi:
                        break
                    if evaluate(case, motion, mid)[3] == s_lo:
                        lo = mid
                    else:
                        hi = mid
                _, _, _, sig_l, m_lo = evaluate(case, motion, lo, probe=True)
                _, _, _, sig_h, m_hi = evaluate(case, motion, hi, probe=True)
                if sig_l == sig_h:
                    continue
                th = A.crossed_threshold(m_lo, m_hi)
                boundaries.append({"i": i, "j": j, "family": family, "motion": motion, "lo": lo, "hi": hi,
                                   "threshold": th[0] if th else "unattributed", "occurrence": th[1] if th else 0,
                                   "margin_lo": th[2] if th else None, "margin_hi": th[3] if th else None})
        print("  pair (%d, %d) %-20s boundaries so far %d  [%.0f s]" % (i, j, family, len(boundaries), time.time() - t_start))
    # keep up to PER_THRESHOLD boundaries per threshold name, preferring different pairs / motions
    by_name = {}
    for b in boundaries:
        by_name.setdefault(b["threshold"], []).append(b)
    kept = []
    for name, lst in sorted(by_name.items()):
        lst.sort(key=lambda b: (b["motion"], b["i"], b["j"]))
        step = max(1, len(lst) // PER_THRESHOLD)
        kept.extend(lst[::step][:PER_THRESHOLD])
    print("boundaries: %d found, %d kept over %d threshold names" % (len(boundaries), len(kept), len(by_name)))
    for name in sorted(by_name):
        print("   %-55s %d" % (name, len(by_name[name])))
    # probe points
    cases = []
    mismatches = 0
    for b in kept:
        case = A.PairCase(spec, b["i"], b["j"])
        points = []
        for side, t0, sgn in (("lo", b["lo"], -1.0), ("hi", b["hi"], 1.0)):
            for d in A.DELTAS:
                t = t0 + sgn * d
                sp, rows, ncand, sig, margins = evaluate(case, b["motion"], t, probe=True)
                ref = reference_rows(sp)
                if ref != rows:
                    mismatches += 1
                    print("ORACLE != REFERENCE at", b, side, d, "margin", A.margin_at(margins, b["threshold"], b["occurrence"]))
                points.append({"t": t, "side": side, "delta": d, "checksum": A.checksum(sp), "rows": ref,
                               "n_candidates": ncand})
        b = dict(b)
        b["points"] = points
        cases.append(b)
    # cube boundaries: the whole pair shifted so that a coordinate of nt1's base centre is 24 -+ d
    cells = []
    for (i, j), family in list(sorted(chosen.items()))[::6]:
        case = A.PairCase(spec, i, j)
        for axis in range(3):
            points = []
            for sgn in (-1.0, 1.0):
                for d in A.DELTAS:
                    shift = np.zeros(3)
                    shift[axis] = 24.0 + sgn * d - case.c1[axis]
                    sp, rows, ncand, sig, margins = evaluate(case, "z", 0.0, shift=shift)
                    ref = reference_rows(sp)
                    if ref != rows:
                        mismatches += 1
                        print("ORACLE != REFERENCE at cell case", i, j, axis, sgn, d)
                    points.append({"shift": [float(x) for x in shift], "delta": sgn * d, "checksum": A.checksum(sp),
                                   "rows": ref, "n_candidates": ncand})
            cells.append({"i": i, "j": j, "family": family, "axis": axis, "points": points})
    out = {"spec": "spec_1gid_full.json.gz", "deltas": A.DELTAS, "cases": cases, "cells": cells,
           "oracle_vs_reference_mismatches": mismatches}
    path = os.path.join(ROOT, "tests", "golden", "adversarial.json.gz")
    with gzip.open(path, "wt") as f:
        json.dump(out, f)
    print("wrote %s: %d boundary cases, %d cell cases, %d points, oracle/reference mismatches %d, %.0f s"
          % (path, len(cases), len(cells), sum(len(c["points"]) for c in cases + cells), mismatches, time.time() - t_start))


if __name__ == "__main__":
    main()
