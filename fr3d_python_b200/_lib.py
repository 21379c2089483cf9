"""ctypes binding of libfr3d_b200.so (C ABI in include/fr3d_b200.h).

There is no CPU fallback: if the shared library is missing or a CUDA device
is not available, every compute entry point raises.
"""

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FR3D_B200_LIB") or os.path.join(_HERE, "lib", "libfr3d_b200.so")   # env: A/B builds
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

ABI_VERSION = 7          # FR3D_ABI_VERSION of include/fr3d_b200.h
BASE_SLOTS = 16
BB_SLOTS = 10
META_INTS = 8
N_PARENTS = 5
MASK_HAS_FRAME = 0x80000000
FLAG_STD_RNA = 1
FLAG_MODIFIED = 2
FLAG_DUP = 4

CAT_SO, CAT_STACKING, CAT_SUGAR_RIBOSE, CAT_BACKBONE, CAT_COPLANAR = 1, 2, 4, 8, 16
SLOT_SO12, SLOT_SO21, SLOT_STACK, SLOT_SR12, SLOT_SR21, SLOT_BPH, SLOT_BR, SLOT_CP, SLOT_BP = range(9)

E_CUDA, E_ARG, E_CAPACITY, E_NOTABLES, E_RANGE = -1, -2, -3, -4, -5

EXPORTS = ["fr3d_version", "fr3d_last_error", "fr3d_init", "fr3d_upload_tables", "fr3d_frames_fit", "fr3d_frames_fit_from",
           "fr3d_pair_search_workspace", "fr3d_pair_search", "fr3d_classify_workspace", "fr3d_classify_pairs", "fr3d_kabsch_batched",
           "fr3d_bond_orientation", "fr3d_radius_search_workspace", "fr3d_radius_search",
           "fr3d_discrepancy_all_vs_all", "fr3d_discrepancy_one_vs_many",
           "fr3d_tail_workspace", "fr3d_annotation_tail", "fr3d_discrepancy_rows_upper", "fr3d_discrepancy_assemble",
           "fr3d_fp64_fma_probe", "fr3d_ingest_last_error", "fr3d_ingest_tables", "fr3d_cif_parse", "fr3d_ingest_sizes",
           "fr3d_ingest_fill", "fr3d_ingest_free",
           "fr3d_order_workspace", "fr3d_tree_penalty", "fr3d_penalized_distance", "fr3d_order_paths", "fr3d_reorder_symmetric",
           "fr3d_possibilities_first", "fr3d_possibilities_extend", "fr3d_frames_fit_ragged", "fr3d_backbone_expand"]
LABEL_COVALENT = 0x40000000
TAIL_MAX_LABELS = 1024
TAIL_MAX_NT = 1 << 17
TAIL_MAX_INDEX = 1 << 23


class Fr3dError(RuntimeError):
    pass


class NtsStruct(C.Structure):
    _fields_ = [("n_nt", C.c_int32), ("reserved", C.c_int32),
                ("center", C.c_void_p), ("rot", C.c_void_p), ("base_xyz", C.c_void_p), ("bb_xyz", C.c_void_p),
                ("base_mask", C.c_void_p), ("bb_mask", C.c_void_p), ("meta", C.c_void_p)]


class StaticTables(C.Structure):
    _fields_ = [("std_xyz", C.c_double * 3 * BASE_SLOTS * N_PARENTS),
                ("n_heavy", C.c_int32 * N_PARENTS), ("n_slots", C.c_int32 * N_PARENTS),
                ("has_div", C.c_int32 * N_PARENTS), ("sr_slot", C.c_int32 * N_PARENTS),
                ("ring_div", C.c_double * 3 * N_PARENTS),
                ("ring5", C.c_double * 3 * 4 * N_PARENTS), ("ring6", C.c_double * 3 * 6 * N_PARENTS),
                ("ell5", C.c_double * 5 * N_PARENTS), ("ell6", C.c_double * 5 * N_PARENTS),
                ("hull", C.c_double * 3 * 7 * N_PARENTS),
                ("n_sites", C.c_int32 * N_PARENTS),
                ("site_h", C.c_int32 * 5 * N_PARENTS), ("site_m", C.c_int32 * 5 * N_PARENTS),
                ("site_carbon", C.c_int32 * 5 * N_PARENTS), ("site_digit", C.c_int32 * 5 * N_PARENTS),
                ("combo_ok", C.c_int32 * (N_PARENTS * N_PARENTS)),
                ("pad_tail", C.c_int32),
                ("so_r2max", C.c_double * N_PARENTS),
                ("hull_r2max", C.c_double * N_PARENTS),
                ("hull_r2in", C.c_double * N_PARENTS)]


SEARCH_MAX_POSITIONS = 16


class SearchQueryStruct(C.Structure):
    _fields_ = [("n_positions", C.c_int32), ("n_units", C.c_int32), ("symbolic", C.c_int32), ("reserved", C.c_int32),
                ("rowptr", C.c_void_p * (SEARCH_MAX_POSITIONS * SEARCH_MAX_POSITIONS)),
                ("cols", C.c_void_p * (SEARCH_MAX_POSITIONS * SEARCH_MAX_POSITIONS)),
                ("universe", C.c_void_p * SEARCH_MAX_POSITIONS), ("centers", C.c_void_p),
                ("perm_dist", C.c_double * (SEARCH_MAX_POSITIONS * SEARCH_MAX_POSITIONS)),
                ("cutoff", C.c_double * SEARCH_MAX_POSITIONS)]


class TailIdentStruct(C.Structure):
    _fields_ = [("n_struct", C.c_int32), ("n_chains", C.c_int32), ("n_ends", C.c_int32), ("max_struct_nt", C.c_int32),
                ("struct_off", C.c_void_p), ("struct_chain_off", C.c_void_p), ("chain_ends_off", C.c_void_p),
                ("nt_chain", C.c_void_p), ("nt_cov", C.c_void_p)]


class TailTablesStruct(C.Structure):
    _fields_ = [("n_labels", C.c_int32), ("n_boxes", C.c_int32), ("labels", C.c_void_p), ("boxes", C.c_void_p),
                ("slot_label", C.c_void_p), ("hydrogen_atoms", C.c_uint64)]


_lock = threading.Lock()
_lib = None


def build(verbose=False, only_if_stale=False):
    """Compile csrc/fr3d_b200.cu for sm_100a into lib/libfr3d_b200.so (in-tree).  Safe when several processes
    (one per GPU) find the library stale at once: a file lock serialises them, the library is written under a
    temporary name and renamed, and a process that waited for the lock skips the compile if it is fresh by then."""
    import fcntl
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if only_if_stale and not needs_build():          # another process compiled it while we waited
                return ""
            tmp = "%s.%d.tmp" % (LIB_PATH, os.getpid())
            cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                   "-shared", "-Xcompiler", "-fPIC", "-cudart", "static", "-o", tmp,
                   os.path.join(CSRC, "fr3d_b200.cu"), os.path.join(CSRC, "ingest.cpp")]
            if verbose:
                cmd.insert(1, "-Xptxas")
                cmd.insert(2, "-v")
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise Fr3dError("nvcc failed:\n" + res.stdout + res.stderr)
            os.replace(tmp, LIB_PATH)
            return res.stdout + res.stderr
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "fr3d_b200.h")]
    return any(os.path.getmtime(s) > t for s in srcs)


def load():
    """dlopen the library and declare prototypes (no CUDA call is made here)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if needs_build() and not os.environ.get("FR3D_B200_LIB"):
            # a fresh checkout, or sources newer than the library (a stale .so would be dlopen'ed silently and e.g.
            # receive a newer table struct): compile with nvcc (sm_100a cross-compiles without a GPU)
            try:
                build(only_if_stale=True)          # another rank may be compiling the same file right now
            except Exception as ex:
                raise Fr3dError("libfr3d_b200.so is missing or older than its sources (%s) and building it failed: %s."
                                "  Run `python -c 'import __graft_entry__ as g; g.build()'`.  There is no CPU fallback."
                                % (LIB_PATH, ex))
        lib = C.CDLL(LIB_PATH)
        missing = [name for name in EXPORTS if not hasattr(lib, name)]
        if missing:
            raise Fr3dError("%s does not export %s: rebuild it (python -c 'import __graft_entry__ as g; g.build()')"
                            % (LIB_PATH, ", ".join(missing)))
        lib.fr3d_version.restype = C.c_int
        if lib.fr3d_version() != ABI_VERSION:
            raise Fr3dError("%s has ABI version %d, this package needs %d: rebuild it"
                            % (LIB_PATH, lib.fr3d_version(), ABI_VERSION))
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        lib.fr3d_version.restype = C.c_int
        lib.fr3d_last_error.restype = C.c_char_p
        lib.fr3d_init.argtypes = [C.c_int]
        lib.fr3d_upload_tables.argtypes = [C.POINTER(StaticTables)]
        lib.fr3d_frames_fit.argtypes = [C.POINTER(NtsStruct), vp, C.c_int, vp]
        lib.fr3d_frames_fit_from.argtypes = [C.POINTER(NtsStruct), vp, i32, vp, vp, C.c_int, vp]
        lib.fr3d_frames_fit_ragged.argtypes = [C.POINTER(NtsStruct), vp, vp, vp, vp, C.c_int, vp]
        lib.fr3d_backbone_expand.argtypes = [vp, vp, i32, vp, vp, i32, vp, vp]
        lib.fr3d_pair_search_workspace.argtypes = [i32]
        lib.fr3d_pair_search_workspace.restype = C.c_size_t
        lib.fr3d_pair_search.argtypes = [C.POINTER(NtsStruct), dbl, vp, C.c_size_t, vp, vp, vp, i64, vp, vp]
        lib.fr3d_classify_workspace.argtypes = [i64, i32]
        lib.fr3d_classify_workspace.restype = C.c_size_t
        lib.fr3d_classify_pairs.argtypes = [C.POINTER(NtsStruct), vp, vp, vp, vp, vp, i64, C.c_uint32, i32, vp,
                                            C.c_size_t, vp, i64, vp, vp]
        lib.fr3d_radius_search_workspace.argtypes = [i32]
        lib.fr3d_radius_search_workspace.restype = C.c_size_t
        lib.fr3d_radius_search.argtypes = [vp, vp, i32, dbl, vp, C.c_size_t, vp, vp, vp, vp, i64, vp, vp]
        lib.fr3d_bond_orientation.argtypes = [C.POINTER(NtsStruct), C.POINTER(C.c_int32), vp, vp, vp]
        lib.fr3d_kabsch_batched.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp]
        lib.fr3d_discrepancy_all_vs_all.argtypes = [vp, vp, vp, i32, i32, dbl, i32, i32, vp, vp]
        lib.fr3d_discrepancy_rows_upper.argtypes = [vp, vp, vp, i32, i32, dbl, i32, i32, vp, vp]
        lib.fr3d_discrepancy_assemble.argtypes = [vp, i32, i32, i32, vp, vp]
        lib.fr3d_fp64_fma_probe.argtypes = [i64, vp, C.POINTER(C.c_int64), vp]
        lib.fr3d_discrepancy_one_vs_many.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, dbl, dbl, vp, vp, vp]
        lib.fr3d_tail_workspace.argtypes = [i64, i32, i32, i64]
        lib.fr3d_tail_workspace.restype = C.c_size_t
        lib.fr3d_annotation_tail.argtypes = [C.POINTER(NtsStruct), C.POINTER(TailIdentStruct), C.POINTER(TailTablesStruct),
                                             vp, i64, vp, i32, vp, C.c_size_t, vp, i64, vp, vp, vp, vp]
        lib.fr3d_order_workspace.argtypes = [i32, i32]
        lib.fr3d_order_workspace.restype = C.c_size_t
        lib.fr3d_tree_penalty.argtypes = [vp, i32, vp, vp, vp, C.c_size_t, vp]
        lib.fr3d_penalized_distance.argtypes = [vp, vp, i32, i32, dbl, vp, vp, vp]
        lib.fr3d_order_paths.argtypes = [vp, i32, vp, i32, i32, vp, vp, vp, vp, C.c_size_t, vp]
        lib.fr3d_reorder_symmetric.argtypes = [vp, i32, vp, i32, vp, vp]
        lib.fr3d_possibilities_first.argtypes = [vp, vp, i64, i32, vp, vp, vp, i64, vp]
        lib.fr3d_possibilities_extend.argtypes = [vp, i32, vp, vp, i64, i32, vp, vp, vp, i64, vp, vp]
        lib.fr3d_ingest_last_error.restype = C.c_char_p
        lib.fr3d_ingest_tables.argtypes = [C.c_char_p, C.c_size_t]
        lib.fr3d_cif_parse.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_int64), i32, i32, C.POINTER(vp)]
        lib.fr3d_ingest_sizes.argtypes = [vp, C.POINTER(C.c_int64)]
        lib.fr3d_ingest_fill.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]
        lib.fr3d_ingest_free.argtypes = [vp]
        lib.fr3d_ingest_free.restype = None
        for name in EXPORTS:
            if name not in ("fr3d_last_error", "fr3d_pair_search_workspace", "fr3d_classify_workspace",
                            "fr3d_radius_search_workspace", "fr3d_tail_workspace", "fr3d_order_workspace", "fr3d_ingest_last_error",
                            "fr3d_ingest_free"):
                getattr(lib, name).restype = C.c_int
        _lib = lib
        return lib


def check(code, what=""):
    if code != 0:
        msg = load().fr3d_last_error().decode(errors="replace")
        raise Fr3dError("%s failed with code %d: %s" % (what or "libfr3d_b200 call", code, msg))
