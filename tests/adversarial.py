"""Adversarial near-threshold cases (north_star: "disagreements are allowed only for pairs within 1e-6 of a rule
cutoff, and each must be listed").

A case is a two-nucleotide structure cut out of 1GID-full whose second nucleotide is moved rigidly along a
one-parameter motion; tools/make_adversarial.py finds, by bisection on the ORACLE's annotation, parameter values t*
at which the annotation of the pair changes, and the probe points are t* -+ {0, 1e-12, 1e-9, 1e-7, 1e-5} on either
side.  Every such point puts one quantity of some rule within that distance (times a slope of order one) of its
cutoff.  oracle.PROBE attributes the boundary to the threshold whose margin changes sign there and gives the
margin at each point.

Shared by the generator (which also asks the unmodified reference for its answer at every point and stores it in
tests/golden/adversarial.json.gz), the CPU test (oracle == reference at every point) and the GPU test (CUDA ==
oracle at every point, disagreements listed with their margins)."""

import numpy as np

from oracle import fr3d_oracle as O

DELTAS = [0.0, 1e-12, 1e-9, 1e-7, 1e-5]
MOTIONS = {           # name -> (t_min, t_max, grid points, unit)
    "z": (-2.6, 2.6, 27, "A along the normal of nt1"),
    "x": (-3.2, 3.2, 27, "A along the x axis of nt1"),
    "y": (-3.2, 3.2, 27, "A along the y axis of nt1"),
    "tilt": (-0.9, 0.9, 25, "rad about an in-plane axis of nt2"),
    "spin": (-1.2, 1.2, 25, "rad about the normal of nt2"),
    "radial": (-3.5, 9.5, 27, "A along the line of base centres"),
}


def _rotate(points, centre, axis, angle):
    """Rodrigues rotation of points [n, 3] about `axis` (unit) through `centre`; element-wise arithmetic only."""
    v = points - centre
    c, s = np.cos(angle), np.sin(angle)
    kx, ky, kz = axis
    cross = np.stack([ky * v[:, 2] - kz * v[:, 1], kz * v[:, 0] - kx * v[:, 2], kx * v[:, 1] - ky * v[:, 0]], axis=1)
    dot = v[:, 0] * kx + v[:, 1] * ky + v[:, 2] * kz
    return centre + v * c + cross * s + np.outer(dot, axis) * (1.0 - c)


class PairCase(object):
    """nt i and nt j of a structure spec; frames of the unmoved pair fix the motion axes."""

    def __init__(self, spec, i, j):
        self.pdb = spec["pdb"]
        self.nt1 = spec["nts"][i]
        self.nt2 = spec["nts"][j]
        self.i, self.j = i, j
        n1, n2 = O.Nt(self.pdb, self.nt1), O.Nt(self.pdb, self.nt2)
        self.c1, self.R1 = np.array(n1.base_center), np.array(n1.rotation)
        self.c2, self.R2 = np.array(n2.base_center), np.array(n2.rotation)
        self.xyz2 = np.array([a[1:4] for a in self.nt2["atoms"]], dtype=float)
        d = self.c2 - self.c1
        self.radial = d / np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])

    def moved(self, motion, t, shift=None):
        """Spec of the pair with nt2 moved by parameter t (and, optionally, both nts shifted by `shift`)."""
        p = self.xyz2
        if motion == "z":
            q = p + t * self.R1[:, 2]
        elif motion == "x":
            q = p + t * self.R1[:, 0]
        elif motion == "y":
            q = p + t * self.R1[:, 1]
        elif motion == "radial":
            q = p + t * self.radial
        elif motion == "tilt":
            q = _rotate(p, self.c2, self.R2[:, 0], t)
        elif motion == "spin":
            q = _rotate(p, self.c2, self.R2[:, 2], t)
        else:
            raise ValueError(motion)
        nt1 = dict(self.nt1)
        nt2 = dict(self.nt2)
        a1 = np.array([a[1:4] for a in self.nt1["atoms"]], dtype=float)
        if shift is not None:
            a1 = a1 + shift
            q = q + shift
        nt1["atoms"] = [[a[0], float(x[0]), float(x[1]), float(x[2])] for a, x in zip(self.nt1["atoms"], a1)]
        nt2["atoms"] = [[a[0], float(x[0]), float(x[1]), float(x[2])] for a, x in zip(self.nt2["atoms"], q)]
        return {"pdb": self.pdb, "nts": [nt1, nt2]}


def annotate_oracle(spec, categories, probe=False):
    """Oracle rows of a spec as {label: [[pos1, pos2, crossing], ...]} without the covalent links (and the list of
    threshold margins when probe is set)."""
    O.PROBE = [] if probe else None
    try:
        out, c2i, tr = O.annotate(spec, categories)
        margins = O.PROBE
    finally:
        O.PROBE = None
    rows = {k: [list(t) for t in v] for k, v in out.items() if v and not k.startswith("p_")}
    return (rows, margins) if probe else rows


def signature(rows):
    return tuple(sorted((k, tuple(map(tuple, v))) for k, v in rows.items()))


def crossed_threshold(m_lo, m_hi):
    """The first threshold whose margin changes sign between the two sides of a boundary: (name, occurrence of that
    name, margin lo, margin hi); None when the two runs diverge before any sign change."""
    seen = {}
    for (na, va), (nb, vb) in zip(m_lo, m_hi):
        if na != nb:
            return None
        k = seen.get(na, 0)
        seen[na] = k + 1
        if (va > 0) != (vb > 0) or (va < 0) != (vb < 0):
            return na, k, va, vb
    return None


def margin_at(margins, name, occurrence):
    k = 0
    for n, v in margins:
        if n == name:
            if k == occurrence:
                return v
            k += 1
    return None


def checksum(spec):
    """A float that pins the coordinates of a generated spec (the golden file stores t, not coordinates)."""
    return float(sum(a[1] + 2.0 * a[2] + 3.0 * a[3] for nt in spec["nts"] for a in nt["atoms"]))


def rows_to_unit_ids(rows, spec):
    nts = O.build_nts(spec)
    return {k: [(nts[a].unit_id(), nts[b].unit_id(), c) for a, b, c in v] for k, v in rows.items()}


# ------------------------------------------------------------------ shared by the tests and tools/adversarial_report.py
def load_cases(golden, spec):
    """Every probe point of the golden file as (key, spec, reference rows, reference candidate count, case, point);
    key = (kind, case index, point index).  The coordinates are regenerated from (pair, motion, t)."""
    out = []
    for ci, c in enumerate(golden["cases"]):
        pc = PairCase(spec, c["i"], c["j"])
        for pi, p in enumerate(c["points"]):
            sp = pc.moved(c["motion"], p["t"])
            sp["pdb"] = "B%03d%02d" % (ci, pi)
            out.append((("boundary", ci, pi), sp, p["rows"], p["n_candidates"], c, p))
    for ci, c in enumerate(golden["cells"]):
        pc = PairCase(spec, c["i"], c["j"])
        for pi, p in enumerate(c["points"]):
            sp = pc.moved("z", 0.0, shift=np.array(p["shift"]))
            sp["pdb"] = "C%03d%02d" % (ci, pi)
            out.append((("cell", ci, pi), sp, p["rows"], p["n_candidates"], c, p))
    return out


def rows_as_positions(i2t, spec):
    """A result dict with unit-id tuples -> {label: [[pos1, pos2, crossing], ...]} without covalent links."""
    pos = {nt.unit_id(): k for k, nt in enumerate(O.build_nts(spec))}
    return {k: [[pos[a], pos[b], c] for a, b, c in v] for k, v in i2t.items() if v and not k.startswith("p_")}


def disagreements(points, got_rows, got_ncand, categories):
    """Compare per point; for every difference report the threshold the boundary belongs to and the margin of that
    threshold at the point (oracle.PROBE).  Returns a list of dicts."""
    out = []
    for (key, sp, want_rows, want_ncand, case, p), rows, ncand in zip(points, got_rows, got_ncand):
        if rows == want_rows and ncand == want_ncand:
            continue
        margin = None
        name = case.get("threshold", "cube boundary")
        if key[0] == "boundary" and name != "unattributed":
            _, margins = annotate_oracle(sp, categories, probe=True)
            margin = margin_at(margins, name, case["occurrence"])
        out.append({"key": key, "pair": (case["i"], case["j"]), "motion": case.get("motion", "shift"), "threshold": name,
                    "delta": p["delta"], "margin": margin,
                    "labels": sorted(set(rows) ^ set(want_rows)) or sorted(k for k in rows if rows[k] != want_rows.get(k)),
                    "candidates": (ncand, want_ncand)})
    return out
