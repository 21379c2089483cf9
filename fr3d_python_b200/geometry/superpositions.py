"""Drop-in fr3d/geometry/superpositions.py (besttransformation :10-89, besttransformation_weighted
:92-121) on the batched CUDA Kabsch kernel (K5), plus the batched forms the GPU wants."""

import ctypes as C

import numpy

from .. import _lib
from ..engine import Engine


def besttransformation_batched(sets1, sets2, weights=None, want_new1=True, device=None):
    """Superpose many problems in one launch.  sets1/sets2: list of (n_k, 3) arrays (or one
    (P, n, 3) array).  weights: None or list of (n_k,) arrays.  Returns dict of numpy arrays:
    U (P,3,3), mean1 (P,3), mean2 (P,3), rmsd (P,), sse (P,), new1 (list of (n_k,3))."""
    eng = Engine.get(device)
    torch = eng.torch
    a1 = [numpy.asarray(s, dtype=numpy.float64).reshape(-1, 3) for s in sets1]
    a2 = [numpy.asarray(s, dtype=numpy.float64).reshape(-1, 3) for s in sets2]
    P = len(a1)
    assert len(a2) == P
    lens = numpy.array([len(s) for s in a1], dtype=numpy.int64)
    for s1, s2 in zip(a1, a2):
        assert len(s1) == len(s2), 'Lengths must match'
        assert len(s1) > 0, 'Must not give empty matrices'
    off = numpy.zeros(P + 1, dtype=numpy.int32)
    off[1:] = numpy.cumsum(lens)
    rows = int(off[-1])
    with torch.cuda.device(eng.device):
        d1 = torch.from_numpy(numpy.concatenate(a1)).to(eng.device)
        d2 = torch.from_numpy(numpy.concatenate(a2)).to(eng.device)
        doff = torch.from_numpy(off).to(eng.device)
        dw = None
        if weights is not None:
            w = numpy.concatenate([numpy.asarray(x, dtype=numpy.float64).reshape(-1) for x in weights])
            assert len(w) == rows
            dw = torch.from_numpy(w).to(eng.device)
        U = torch.empty((P, 9), dtype=torch.float64, device=eng.device)
        m1 = torch.empty((P, 3), dtype=torch.float64, device=eng.device)
        m2 = torch.empty((P, 3), dtype=torch.float64, device=eng.device)
        rm = torch.empty(P, dtype=torch.float64, device=eng.device)
        ss = torch.empty(P, dtype=torch.float64, device=eng.device)
        n1 = torch.empty((rows, 3), dtype=torch.float64, device=eng.device) if want_new1 else None
        _lib.check(eng.lib.fr3d_kabsch_batched(d1.data_ptr(), d2.data_ptr(), dw.data_ptr() if dw is not None else None,
                                               doff.data_ptr(), P, U.data_ptr(), m1.data_ptr(), m2.data_ptr(),
                                               rm.data_ptr(), ss.data_ptr(),
                                               n1.data_ptr() if n1 is not None else None, eng.stream_ptr()),
                   "fr3d_kabsch_batched")
        eng.launches += 1
        out = {"U": U.cpu().numpy().reshape(P, 3, 3), "mean1": m1.cpu().numpy(), "mean2": m2.cpu().numpy(),
               "rmsd": rm.cpu().numpy(), "sse": ss.cpu().numpy()}
        if want_new1:
            flat = n1.cpu().numpy()
            out["new1"] = [flat[off[k]:off[k + 1]] for k in range(P)]
    return out


def besttransformation(set1, set2):
    """superpositions.py:10-89 -> (U, new1, mean1, rmsd, sse, mean2); U and new1 are numpy.matrix
    like the reference."""
    assert len(set1) == len(set2), 'Lengths must match'
    assert len(set1) > 0, 'Must not give empty matrices'
    r = besttransformation_batched([set1], [set2])
    return (numpy.matrix(r["U"][0]), numpy.matrix(r["new1"][0]), r["mean1"][0], float(r["rmsd"][0]),
            float(r["sse"][0]), r["mean2"][0])


def besttransformation_weighted(set1, set2, weights=[1.0]):
    """superpositions.py:92-121 -> (U, new1, mean1, rmsd, sse); weights are ignored unless there is
    one per point (:101-104); sse / rmsd are unweighted (RMSD.py:28)."""
    assert len(set1) == len(set2)
    assert len(set1) > 0
    w = [weights] if len(weights) == len(set1) else None
    r = besttransformation_batched([set1], [set2], w)
    return (numpy.matrix(r["U"][0]), numpy.matrix(r["new1"][0]), r["mean1"][0], float(r["rmsd"][0]),
            float(r["sse"][0]))
