"""The C-ABI library loads and exports exactly what include/fr3d_b200.h declares; ctypes mirrors of
the structs have the C compiler's sizes.  No compute call is made (no GPU needed)."""

import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np

from fr3d_python_b200 import _lib
from fr3d_python_b200 import definitions as defs
from fr3d_python_b200.engine import BOX_DTYPE, HIT_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fr3d_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fr3d_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    if _lib.needs_build():
        _lib.build()
    lib = _lib.load()
    names = declared_functions()
    assert set(names) == set(_lib.EXPORTS), (names, _lib.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.fr3d_version() == 1
    assert lib.fr3d_last_error() is not None
    assert lib.fr3d_pair_search_workspace(1000) > 0
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (fr3d_[a-z_0-9]+)", out))
    assert exported == set(names)


def test_struct_sizes_match_the_c_compiler(tmp_path):
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "fr3d_b200.h"\nint main(void){printf("%zu %zu %zu %zu\\n", '
                   'sizeof(fr3d_hit_t), sizeof(fr3d_nts_t), sizeof(fr3d_static_tables_t), sizeof(fr3d_box_t));return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    assert sizes == [HIT_DTYPE.itemsize, C.sizeof(_lib.NtsStruct), C.sizeof(_lib.StaticTables), BOX_DTYPE.itemsize]
    assert sizes[0] == 32 and sizes[3] == 128


def test_static_tables_content():
    t = defs.static_tables()
    assert list(t.n_heavy) == [10, 8, 11, 8, 9] and list(t.n_slots) == [14, 12, 15, 11, 14]
    # slot 0 is the glycosidic atom, standard bases are centred on their heavy atoms
    for p, parent in enumerate(defs.PARENTS):
        assert defs.SLOT_NAMES[parent][0] in ("N9", "N1")
        xyz = np.array([[t.std_xyz[p][k][c] for c in range(3)] for k in range(t.n_heavy[p])])
        assert np.abs(xyz.mean(axis=0)).max() < 1e-6 and np.all(xyz[:, 2] == 0)
    assert sum(t.combo_ok) == 14
    assert list(t.has_div) == [1, 0, 1, 0, 0]
    assert [t.site_digit[2][k] for k in range(t.n_sites[2])] == [5, 0, 1, 3]      # G: 5BPh 0BPh 1BPh 3BPh


def test_header_cites_reference_lines():
    text = open(HEADER).read()
    for needle in ("components.py:273-349", "NA:410-443", "NA:1278", "superpositions.py:10-121",
                   "discrepancy.py:165-233", "FR3D.py:433-442"):
        assert needle in text
