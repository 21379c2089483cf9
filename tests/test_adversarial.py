"""Near-threshold parity (north_star: "disagreements are allowed only for pairs within 1e-6 of a rule cutoff, and
each must be listed").  tests/golden/adversarial.json.gz holds 220 decision boundaries of the rules, found by
bisection (tools/make_adversarial.py) and attributed to 47 named thresholds - stacking min distance 1 / 4, |normal_Z|
0.5 / 0.6, hull lines, |z| 4.5; sO |z| 2 / 3.5 / 3.6, ring lines, ellipses; BPh / BR angles 110 / 130 and distances;
SR 3.8 / 2.0; the coplanar thresholds; cutoff-box edges, angles, gap and the cutoff_distance 1.0 line; the 2 / 12 A
search screens - plus pairs shifted onto cube boundaries, each probed at +-{0, 1e-12, 1e-9, 1e-7, 1e-5} around the
boundary, with the answer of the UNMODIFIED reference at every point."""

import os
import subprocess
import sys

import numpy as np
import pytest

import adversarial as A
from conftest import ALL9, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _points():
    golden = load_golden("adversarial.json.gz")
    spec = load_golden(golden["spec"])
    return golden, A.load_cases(golden, spec)


def test_golden_covers_the_rule_thresholds():
    golden, points = _points()
    names = set(c["threshold"] for c in golden["cases"])
    for needle in ("stacking min distance vs 1", "stacking min distance vs 4", "stacking |normal_Z| vs 0.5",
                   "stacking |normal_Z| vs 0.6", "stacking hull line", "stacking atom |z| vs 4.5", "sO |z| vs 2",
                   "sO |z| vs 3.5", "sO |z| vs 3.6", "sO ring6 line", "sO near-ring ellipse", "BPh angle vs 110",
                   "BPh angle vs 130", "BR angle vs 130", "BPh distance vs 4.0 (C) / 3.5 (N)", "SR O2'-O2' vs 3.8",
                   "SR height vs 2.0", "coplanar gap12 vs 1.5179", "coplanar |c.n1| vs 0.3381",
                   "coplanar |n1.n2| vs 0.7757", "coplanar min distance vs 2.4589", "basepair box edge xmin",
                   "basepair box angle min", "basepair gap vs gapmax", "basepair cutoff_distance vs 1.0 (with gap)",
                   "basepair |z| vs 3.6", "search d vs 12", "search d vs 2"):
        assert needle in names, needle
    assert len(points) >= 2000 and golden["oracle_vs_reference_mismatches"] == 0
    assert sorted(set(abs(p["delta"]) for c in golden["cases"] for p in c["points"])) == [0.0, 1e-12, 1e-9, 1e-7, 1e-5]


def test_oracle_equals_reference_at_every_probe_point():
    """Pins the oracle where it matters most: at every probe point the oracle's rows and candidate count equal what
    the unmodified reference produced (the coordinates are regenerated here; the checksum proves they are the same)."""
    golden, points = _points()
    bad = []
    for key, sp, want_rows, want_ncand, case, p in points:
        assert A.checksum(sp) == p["checksum"], key
        out, c2i, tr = A.O.annotate(sp, ALL9)
        rows = {k: [list(t) for t in v] for k, v in out.items() if v and not k.startswith("p_")}
        if rows != want_rows or len(tr["candidates"]) != want_ncand:
            bad.append(key)
    assert not bad, bad


_SUBPROCESS = r"""
import json, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
import adversarial as A
from conftest import ALL9, load_golden
from fr3d_python_b200 import packing
from fr3d_python_b200.classifiers import NA_pairwise_interactions as NA
golden = load_golden("adversarial.json.gz")
points = A.load_cases(golden, load_golden(golden["spec"]))
specs = [p[1] for p in points]
batch = packing.pack_specs(specs)
res, ex = NA.annotate_batch(batch, ALL9, return_raw=True)
sid = np.searchsorted(batch.struct_off, ex["pairs"][0], side="right") - 1
ncand = np.bincount(sid, minlength=len(specs)).tolist()
rows = [A.rows_as_positions(res[s][0], specs[s]) for s in range(len(specs))]
json.dump({"rows": rows, "ncand": ncand}, open(sys.argv[1], "w"))
"""


@pytest.mark.gpu
@pytest.mark.parametrize("screen", ["32", "64"])
def test_cuda_at_the_rule_thresholds(screen, tmp_path):
    """All probe points in ONE batch through annotate_batch (default path: fp32 screen + fp32-certain stacking;
    FR3D_SCREEN=64: the all-fp64 kernels).  Every difference from the reference's answer is listed with the margin
    of its threshold at that point, which must be within 1e-6."""
    import json
    golden, points = _points()
    out = str(tmp_path / ("adv_%s.json" % screen))
    code = _SUBPROCESS % (ROOT, os.path.join(ROOT, "tests"))
    subprocess.run([sys.executable, "-c", code, out], check=True, env=dict(os.environ, FR3D_SCREEN=screen), cwd=ROOT)
    with open(out) as f:
        got = json.load(f)
    diffs = A.disagreements(points, got["rows"], got["ncand"], ALL9)
    report = ["%d probe points, %d boundaries over %d thresholds, FR3D_SCREEN=%s: %d disagreements"
              % (len(points), len(golden["cases"]), len(set(c["threshold"] for c in golden["cases"])), screen, len(diffs))]
    for d in diffs:
        report.append("  %-9s pair %-10s %-7s %-50s delta %-8g margin %s labels %s candidates %s"
                      % (d["key"][0], d["pair"], d["motion"], d["threshold"], d["delta"],
                         "%.3e" % d["margin"] if d["margin"] is not None else "n/a", d["labels"], d["candidates"]))
    print("\n".join(report))
    dest = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(dest):
        with open(os.path.join(dest, "adversarial_report_screen%s.txt" % screen), "w") as f:
            f.write("\n".join(report) + "\n")
    for d in diffs:
        if d["margin"] is not None:
            assert abs(d["margin"]) <= 1e-6, d
        else:
            assert abs(d["delta"]) <= 1e-9, d            # unattributed boundary / cube boundary: the parameter offset itself
    assert len(diffs) <= 0.02 * len(points)
