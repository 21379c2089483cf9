"""The C-ABI library loads and exports exactly what include/fr3d_b200.h declares; ctypes mirrors of
the structs have the C compiler's sizes.  No compute call is made (no GPU needed)."""

import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

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
    header_version = int(re.search(r"#define FR3D_ABI_VERSION (\d+)", open(HEADER).read()).group(1))
    assert lib.fr3d_version() == _lib.ABI_VERSION == header_version
    assert lib.fr3d_last_error() is not None          # (the driver runs __graft_entry__.build() itself each round)
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


@pytest.mark.parametrize("per_parent", [False, True])
def test_rings_and_ellipses_lie_inside_the_prefilter_disc(per_parent):
    """csrc/classify_tpp.cuh skips the ring / ellipse tests for oxygens with x^2 + y^2 >= 25 and the screen kernel
    uses the per-parent radius definitions.SCREEN_SO_RADIUS: every ring polygon and near-ring ellipse (convex)
    must lie strictly inside that disc.  A convex region with a point inside the disc and no point on the
    circle is contained in the disc."""
    ang = np.linspace(0, 2 * np.pi, 7200, endpoint=False)

    def inside(lines, x, y):
        ok = np.ones_like(x, dtype=bool)
        for a, b, c in lines:
            ok &= a * x + b * y + c > 0
        return ok

    def in_ell(e, x, y):
        a, b, ex, ey, r = e
        return a * (x - ex) ** 2 + b * (x - ex) * (y - ey) + (y - ey) ** 2 < r

    for parent in defs.PARENTS:
        radius = defs.SCREEN_SO_RADIUS[parent] if per_parent else 5.0
        cx, cy = radius * np.cos(ang), radius * np.sin(ang)
        rings = defs.PLANES["rings"][parent]
        ell = defs.PLANES["ellipses"][parent]
        regions = []
        if rings["divider"]:
            d = rings["divider"]
            regions.append([d] + rings["ring5"])
            regions.append([[-d[0], -d[1], -d[2] + 1e-12]] + rings["ring6"])
        else:
            regions.append(rings["ring6"])
        gx, gy = np.meshgrid(np.linspace(-4, 4, 161), np.linspace(-4, 4, 161))
        for reg in regions:
            assert not inside(reg, cx, cy).any(), parent
            assert inside(reg, gx.ravel(), gy.ravel()).any(), parent        # non-empty, near the origin
        for e in (ell["near5"], ell["near6"]):
            if e:
                assert not in_ell(e, cx, cy).any(), parent
                assert in_ell(e, np.array([e[2]]), np.array([e[3]])).all()


@pytest.mark.parametrize("per_parent", [False, True])
def test_hulls_lie_inside_the_screen_disc(per_parent):
    """The thread-per-pair kernels test the hull half planes only for atoms with x^2 + y^2 < 16, the screen kernel
    uses the per-parent radius definitions.SCREEN_HULL_RADIUS: every convex hull of check_convex_hull_atoms
    (NA:1485-1523) must lie strictly inside that disc."""
    ang = np.linspace(0, 2 * np.pi, 7200, endpoint=False)
    for parent in defs.PARENTS:
        radius = defs.SCREEN_HULL_RADIUS[parent] if per_parent else 4.0
        cx, cy = radius * np.cos(ang), radius * np.sin(ang)
        ok = np.ones_like(cx, dtype=bool)
        inside0 = True
        for a, b, c in defs.PLANES["hull"][parent]:
            ok &= a * cx + b * cy + c > 0
            inside0 = inside0 and c > 0
        assert not ok.any(), parent
        assert inside0, parent                      # the base centre is inside the hull


def test_inner_disc_lies_inside_the_hull():
    """The overlap tests accept an atom with x^2 + y^2 < definitions.HULL_INNER_RADIUS[parent]^2 without
    evaluating the half planes: that disc must lie strictly inside the convex hull (every point of its boundary
    circle, and therefore the whole disc, satisfies all half planes with room to spare)."""
    ang = np.linspace(0, 2 * np.pi, 7200, endpoint=False)
    for parent in defs.PARENTS:
        radius = defs.HULL_INNER_RADIUS[parent] + 0.01
        cx, cy = radius * np.cos(ang), radius * np.sin(ang)
        for a, b, c in defs.PLANES["hull"][parent]:
            assert (a * cx + b * cy + c > 0).all(), parent
