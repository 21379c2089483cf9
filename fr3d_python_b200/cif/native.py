"""Native mmCIF ingest: CIF texts -> NtBatch through the C++ reader of libfr3d_b200.so (csrc/ingest.cpp), one thread
per file.  Same result as cif.reader.cif_to_batch (the Python statement of the same logic; tests/test_cif_ingest.py
compares every array and identity column); the record planes can be written straight into pinned memory."""

import ctypes as C

import numpy as np

from .. import _lib
from .. import definitions as defs
from .. import packing

_tables_loaded = False


def _tables_blob():
    """The slot maps of every residue type packing._slot_maps knows, as the text table fr3d_ingest_tables reads."""
    lines = []
    seqs = list(defs.SEQ_TO_PARENT.keys()) + sorted(defs.modified_table().keys())
    for seq in seqs:
        m = packing._slot_maps(seq)
        if m is None:
            continue
        pcode, flags, base, bb, dup = m
        lines.append("S\t%s\t%d\t%d\t%d" % (seq, pcode, flags, dup))
        for nm, slot in base.items():
            lines.append("B\t%s\t%d" % (nm, slot))
        for nm, slot in bb.items():
            lines.append("K\t%s\t%d" % (nm, slot))
    for parent, (heavy, h1, h2) in packing._AMINO.items():
        so = defs.SLOT_OF[parent]
        lines.append("A\t%d\t%d\t%d\t%d" % (defs.PARENT_CODE[parent], so[heavy], so[h1], so[h2]))
    return ("\n".join(lines) + "\n").encode()


def _check(lib, code, what):
    if code != 0:
        raise _lib.Fr3dError("%s failed with code %d: %s" % (what, code, lib.fr3d_ingest_last_error().decode(errors="replace")))


def _column(buf):
    vals = buf.tobytes().decode().split("\0")[:-1]
    return [None if v == "\1" else v for v in vals]


def cif_to_batch(texts, n_threads=0, pinned=False):
    """NtBatch of the nucleotides of a list of CIF texts (str or bytes).  pinned=True: the record planes live in pinned
    host memory (torch), ready for engine.pin_batch-free uploads."""
    global _tables_loaded
    if isinstance(texts, (str, bytes)):
        texts = [texts]
    lib = _lib.load()
    if not _tables_loaded:
        blob = _tables_blob()
        _check(lib, lib.fr3d_ingest_tables(blob, len(blob)), "fr3d_ingest_tables")
        _tables_loaded = True
    raw = [t.encode() if isinstance(t, str) else bytes(t) for t in texts]
    arr = (C.c_char_p * len(raw))(*raw)
    lens = (C.c_int64 * len(raw))(*[len(r) for r in raw])
    handle = C.c_void_p()
    _check(lib, lib.fr3d_cif_parse(arr, lens, len(raw), int(n_threads), C.byref(handle)), "fr3d_cif_parse")
    try:
        sizes = (C.c_int64 * 9)()
        _check(lib, lib.fr3d_ingest_sizes(handle, sizes), "fr3d_ingest_sizes")
        n, S = int(sizes[0]), int(sizes[1])
        b = packing.NtBatch(n)
        if pinned:
            import torch
            for name in ("base_xyz", "bb_xyz", "base_mask", "bb_mask", "meta"):
                a = getattr(b, name)
                t = torch.zeros(a.shape, dtype={np.dtype('float64'): torch.float64, np.dtype('uint32'): torch.int32,
                                                np.dtype('int32'): torch.int32}[a.dtype]).pin_memory()
                setattr(b, name, t.numpy().view(a.dtype))
                setattr(b, "_pin_" + name, t)
        struct_off = np.zeros(S + 1, dtype=np.int64)
        number = np.zeros(max(n, 1), dtype=np.int64)
        index = np.zeros(max(n, 1), dtype=np.int64)
        cols = [np.zeros(max(int(sizes[2 + k]), 1), dtype=np.uint8) for k in range(7)]
        col_ptrs = (C.c_void_p * 7)(*[c.ctypes.data for c in cols])
        _check(lib, lib.fr3d_ingest_fill(handle, b.base_xyz.ctypes.data, b.bb_xyz.ctypes.data, b.base_mask.ctypes.data,
                                         b.bb_mask.ctypes.data, b.meta.ctypes.data, struct_off.ctypes.data,
                                         number.ctypes.data, index.ctypes.data, col_ptrs), "fr3d_ingest_fill")
    finally:
        lib.fr3d_ingest_free(handle)
    b.struct_off = struct_off
    b.model, b.chain, b.sequence, b.symmetry, b.alt_id, b.insertion_code, b.pdb = \
        [_column(c[:int(sizes[2 + k])]) for k, c in enumerate(cols)]
    b.number = number[:n].tolist()
    imin = np.iinfo(np.int64).min
    b.index = [None if v == imin else v for v in index[:n].tolist()]
    return b
