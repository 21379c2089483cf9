/*
 * fr3d_b200.h — C ABI of libfr3d_b200.so (sm_100a).
 *
 * The reference (BGSU-RNA/fr3d-python) is pure Python and has no FFI; the
 * boundary it exposes is a set of Python call signatures.
 * Each entry point below names the reference function(s) whose arithmetic it
 * replaces; fr3d_python_b200/ keeps the Python signatures on top of this ABI
 * through ctypes.
 *
 * Conventions
 *  - every function returns 0 (FR3D_OK) or a negative FR3D_E_* code, never
 *    throws, never prints; fr3d_last_error() gives a thread-local message.
 *  - all data pointers are DEVICE pointers owned by the caller (torch
 *    tensors' data_ptr()) unless the parameter name starts with host_.
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *  - the library allocates nothing that outlives a call except the constant
 *    tables uploaded by fr3d_upload_tables.
 *  - all floating point in the interface is IEEE fp64, and everything that reaches an output is decided in
 *    fp64 (single precision is used internally only for superset screens and for comparisons that are certain
 *    by a margin far above its rounding error; see fr3d_classify_pairs).
 */
#ifndef FR3D_B200_H
#define FR3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FR3D_ABI_VERSION 7

#define FR3D_OK 0
#define FR3D_E_CUDA (-1)      /* a CUDA runtime call failed; see fr3d_last_error */
#define FR3D_E_ARG (-2)       /* bad argument */
#define FR3D_E_CAPACITY (-3)  /* output buffer too small; required count left in the counter */
#define FR3D_E_NOTABLES (-4)  /* fr3d_upload_tables has not been called */
#define FR3D_E_RANGE (-5)     /* coordinates outside the spatial-hash key range */

/* ---- layout constants ---------------------------------------------------- */
#define FR3D_BASE_SLOTS 16 /* base atoms per nt: heavy atoms then hydrogens of the parent base */
#define FR3D_BB_SLOTS 10   /* P OP1 OP2 O5' O4' O3' O2' C1' C2' prevO3' */
#define FR3D_META_INTS 8   /* parent, flags, group, chain_rank, index, resid_key, alt_rank, dup/twin masks */
#define FR3D_N_PARENTS 5   /* A C G U DT */

#define FR3D_MASK_HAS_FRAME 0x80000000u /* bit 31 of base_mask: centre + rotation valid */

/* meta[1] flags */
#define FR3D_FLAG_STD_RNA 1  /* sequence is one of A C G U   (check_coplanar, NA:2095) */
#define FR3D_FLAG_MODIFIED 2 /* modified residue: centre = meanR - U.meanS (components.py:348) */
#define FR3D_FLAG_DUP 4      /* modified residue without observed hydrogens: the reference appends an ideal-position
                                copy of every mapped base atom (components.py:505-513, mapping.py:76-79).
                                meta[7] bits 0-15: slots that receive an ideal copy (input of fr3d_frames_fit);
                                bits 16-31: "twin" slots holding BOTH an observed atom and its ideal copy, stored
                                as their mean (= what AtomProxy returns); rules that loop over atoms() instead of
                                centers[] (gap, minimum distance) recover both points on the device */

/* category bits for fr3d_classify_pairs (keys of the reference's `categories` dict) */
#define FR3D_CAT_SO 1
#define FR3D_CAT_STACKING 2
#define FR3D_CAT_SUGAR_RIBOSE 4
#define FR3D_CAT_BACKBONE 8
#define FR3D_CAT_COPLANAR 16

/* hit slots, in the order the reference evaluates them for one candidate pair
 * (NA_pairwise_interactions.py:758-1013) */
enum {
    FR3D_SLOT_SO12 = 0,  /* code: oxygen(0..5 = O2' O3' O4' O5' OP1 OP2) | face5<<3 | near<<4 */
    FR3D_SLOT_SO21 = 1,  /* same, base of nt2 / oxygen of nt1 */
    FR3D_SLOT_STACK = 2, /* code: 0 s35, 1 s53, 2 s33, 3 s55 | near<<2 */
    FR3D_SLOT_SR12 = 3,  /* code: 0 cSR, 1 tSR */
    FR3D_SLOT_SR21 = 4,
    FR3D_SLOT_BPH = 5,   /* code: digit | near<<4   ("7BPh", "n7BPh") */
    FR3D_SLOT_BR = 6,    /* code: digit | near<<4 */
    FR3D_SLOT_CP = 7,
    FR3D_SLOT_BP = 8     /* code: bit0 near ("n" prefix), bit1 orientation 21 was used; box = cutoff box */
};

/* One annotation found for one candidate pair. 32 bytes. */
typedef struct {
    int32_t i;              /* nt1 of the candidate pair (global nt index) */
    int32_t j;              /* nt2 of the candidate pair */
    int32_t rank;           /* first_seen(cell of nt1) * 16 + stencil index: reference visit order */
    uint8_t slot;           /* FR3D_SLOT_* */
    uint8_t code;           /* label code, meaning depends on slot */
    uint16_t box;           /* FR3D_SLOT_BP: index into the uploaded box table */
    double cutoff_distance; /* FR3D_SLOT_BP: quality['cutoff_distance'] (NA:2347) */
    double max_gap;         /* FR3D_SLOT_BP: quality['max_gap']         (NA:2348) */
} fr3d_hit_t;

/* Nucleotide records, structure-of-arrays, n_nt entries each (device pointers).
 * Flattens what the annotator reads from fr3d.data.Component
 * (fr3d/data/components.py:118-202, fr3d/data/base.py:80-209). */
typedef struct {
    int32_t n_nt;
    int32_t reserved;
    double *center;      /* [n][3]      centers['base'] */
    double *rot;         /* [n][9]      rotation_matrix, row major */
    double *base_xyz;    /* [n][16][3]  base atoms by parent slot */
    double *bb_xyz;      /* [n][10][3]  backbone atoms */
    uint32_t *base_mask; /* [n] bits 0-15: slot k present; 16-19: parent + 1; 20-23: meta flags; 31: frame valid */
    uint32_t *bb_mask;   /* [n] bit k: backbone slot k present */
    int32_t *meta;       /* [n][8] */
} fr3d_nts_t;

/* Constant rule tables (HOST struct, copied to __constant__ memory). Built by the
 * Python layer from fr3d_python_b200/tables/{definitions,planes}.json. */
typedef struct {
    double std_xyz[FR3D_N_PARENTS][FR3D_BASE_SLOTS][3]; /* definitions.py:467-611 */
    int32_t n_heavy[FR3D_N_PARENTS];
    int32_t n_slots[FR3D_N_PARENTS];
    int32_t has_div[FR3D_N_PARENTS];
    int32_t sr_slot[FR3D_N_PARENTS];       /* N3 (A,G) / O2 (C,U) slot, -1 none (NA:2291-2296) */
    double ring_div[FR3D_N_PARENTS][3];    /* NA:1318,1340 */
    double ring5[FR3D_N_PARENTS][4][3];    /* NA:1319-1322 ... */
    double ring6[FR3D_N_PARENTS][6][3];    /* padded with {0,0,1} */
    double ell5[FR3D_N_PARENTS][5];        /* a b cx cy r  (NA:1415,1425) */
    double ell6[FR3D_N_PARENTS][5];
    double hull[FR3D_N_PARENTS][7][3];     /* NA:1485-1523, padded with {0,0,1} */
    int32_t n_sites[FR3D_N_PARENTS];       /* definitions.py:260-298 */
    int32_t site_h[FR3D_N_PARENTS][5];     /* hydrogen slot */
    int32_t site_m[FR3D_N_PARENTS][5];     /* massive atom slot */
    int32_t site_carbon[FR3D_N_PARENTS][5];/* 1: "C" in name, 0: "N" in name, -1 neither */
    int32_t site_digit[FR3D_N_PARENTS][5]; /* n of "nBPh"/"nBR" */
    int32_t combo_ok[FR3D_N_PARENTS * FR3D_N_PARENTS]; /* NA:654 */
    int32_t pad_tail;
    /* Screen discs (classify_screen.cuh), squared radii about the base centre: every ring and near-ring ellipse of
     * the parent lies inside so_r2max, its convex hull inside hull_r2max (proved in tests/test_abi.py). */
    double so_r2max[FR3D_N_PARENTS];
    double hull_r2max[FR3D_N_PARENTS];
    double hull_r2in[FR3D_N_PARENTS];      /* ... and the disc of this squared radius lies inside the hull */
} fr3d_static_tables_t;

/* One basepair cutoff box (class_limits_2023.py after focus_basepair_cutoffs, NA:88-136). 128 B. */
typedef struct {
    double xmin, xmax, ymin, ymax, zmin, zmax, normalmin, normalmax, anglemin, anglemax, gapmax, radiusmax;
    int32_t name_id;      /* host-side id of the interaction name */
    int32_t lower_id;     /* id of name.lower() */
    int32_t rev_lower_id; /* id of reverse_edges(name).lower() */
    int32_t subcat;
    int32_t name_len;
    int32_t flags;        /* bit0 name starts with 'n', bit1 name=='cSs', bit2 name=='csS', bit3 has radiusmax */
    int32_t pad[2];
} fr3d_box_t;

/* ---- library ------------------------------------------------------------ */
int fr3d_version(void);
const char *fr3d_last_error(void);
int fr3d_init(int device);   /* REQUIRED once per process and device before any other call: checks sm_100 and opts
                                  the kernels in to their shared-memory carve-outs on that device.  The caller's
                                  current device is left unchanged.  State is kept per device, so one thread per GPU
                                  in one process works; every other entry point acts on the CURRENT device and fails
                                  with FR3D_E_ARG if that device was not initialised. */
int fr3d_upload_tables(const fr3d_static_tables_t *host_tables);   /* to the current device (FR3D_E_NOTABLES without) */

/* K1. Replaces Component.calculate_rotation_matrix (fr3d/data/components.py:273-349) +
 * besttransformation (fr3d/geometry/superpositions.py:10-89) for every nt in one launch, and
 * the hydrogen placement arithmetic of infer_NA_hydrogens (components.py:452-462) when
 * place_h != 0.  Reads base_xyz/base_mask/meta; writes center, rot, missing hydrogen slots,
 * base_mask (hydrogen bits + FR3D_MASK_HAS_FRAME) and status[n] (0 ok, -1 no parent, -2 <3 atoms). */
/* The same from a RAGGED upload plane: ragged_xyz holds only the observed atoms of slots 0..10 (every heavy base atom), nt by
 * nt in slot order; group_off[g] = first atom of the 32 nts of group g ([ceil(n_nt / 32) + 1]); observed_mask as above.
 * ~9.3 instead of 11 (or 16) slots per nt cross PCIe; the full 16-slot records are written to nts->base_xyz. */
int fr3d_frames_fit_ragged(const fr3d_nts_t *nts, const double *ragged_xyz, const int32_t *group_off,
                           const uint32_t *observed_mask, int32_t *status, int place_h, void *stream);
/* Companion for the backbone plane: bb9 [n_nt][9][3] (slots 0..8) -> bb_xyz [n_nt][10][3]; slot 9, the O3' of the previous
 * residue (NA_pairwise_interactions.py:480-525), is the previous ROW's O3' where prev_rel[n] == 1 and comes from the explicit
 * list (extra_nt [n_extra], extra_xyz [n_extra][3]) otherwise; the presence bit lives in nts->bb_mask as always. */
int fr3d_backbone_expand(const double *bb9, const int8_t *prev_rel, int32_t n_nt, const int32_t *extra_nt,
                         const double *extra_xyz, int32_t n_extra, double *bb_xyz, void *stream);
int fr3d_frames_fit(const fr3d_nts_t *nts, int32_t *status, int place_h, void *stream);
/* Same, with the observed base atoms read from a compact plane [n][compact_slots][3] (11 <= compact_slots <= 16:
 * the first slots of every record; an upload without atoms in slots 11..15 keeps slots 0..10 only) and their
 * presence bits from observed_mask [n] (same bit layout as base_mask; a plane of its own, never written) instead of
 * nts->base_xyz / nts->base_mask, to which the full 16-slot records and the completed masks are written: a pure
 * function of the upload, repeatable.  compact_base_xyz == NULL: fr3d_frames_fit. */
int fr3d_frames_fit_from(const fr3d_nts_t *nts, const double *compact_base_xyz, int32_t compact_slots,
                         const uint32_t *observed_mask, int32_t *status, int place_h, void *stream);

/* K2+K3. Replaces make_nt_cubes_half (NA:410-443) and the loop/screens of
 * annotate_nt_nt_interactions (NA:661-724): oriented candidate pairs with 2 <= d <= 12.
 * counters (device int64[8]): [0] n_pairs (may exceed cap: FR3D_E_CAPACITY is NOT raised here,
 * the caller compares after its sync), [1] n_hits, [2] error flags, [3] n_cells, [4] [5] [6] [7] work-list sizes of
 * fr3d_classify_pairs (sO + sugar-ribose, stacking, coplanar / basepair, base-phosphate). */
size_t fr3d_pair_search_workspace(int32_t n_nt);
int fr3d_pair_search(const fr3d_nts_t *nts, double cell_size, void *workspace, size_t workspace_bytes,
                     int32_t *pair_i, int32_t *pair_j, int32_t *pair_rank, int64_t pair_cap,
                     int64_t *counters, void *stream);

/* K4. Replaces check_base_oxygen_stack_rings (NA:1278), check_base_base_stacking (NA:1602),
 * check_sugar_ribose (NA:2262), check_base_backbone_interactions (NA:1867), check_coplanar
 * (NA:2051), check_basepair_cutoffs (NA:2364) and the 12-vs-21 arbitration (NA:935-957):
 * hits compacted with a warp-aggregated atomic.
 * n_pairs is read from counters[0] on the device (clamped to pair_cap).  index_offset is added to the
 * nt indices (and 16 x to the rank) written into the hit records, so that chunks of a larger batch
 * produce globally indexed hits.  workspace (fr3d_classify_workspace(pair_cap, n_nt) bytes, may be NULL): work
 * lists of the screen kernel and its single-precision copy of the records (320 B per nt); with it the call is
 * six launches: the screen records, a thread-per-pair single-precision screen of the cheap necessary conditions
 * of every rule family (thresholds relaxed by margins far above fp32 rounding, so it is a superset), then one
 * thread-per-pair fp64 launch per work list (sO / sugar-ribose; BPh / BR; stacking; coplanar / basepair) over the
 * pairs that passed; without it one warp-per-pair launch over all pairs.  Every label and number in a hit record
 * is decided in fp64.  Base centres must lie within +-32 000 A (else counters[2] != 0, FR3D_E_RANGE). */
size_t fr3d_classify_workspace(int64_t pair_cap, int32_t n_nt);
int fr3d_classify_pairs(const fr3d_nts_t *nts, const fr3d_box_t *boxes, const int32_t *box_index /*[25][2][2]*/,
                        const int32_t *pair_i, const int32_t *pair_j, const int32_t *pair_rank,
                        int64_t pair_cap, uint32_t category_mask, int32_t index_offset, void *workspace,
                        size_t workspace_bytes, fr3d_hit_t *hits, int64_t hit_cap, int64_t *counters, void *stream);

/* A8 on the device.  Replaces, for every structure of the batch in one call, the order-dependent
 * tail of annotate_nt_nt_interactions: the visit-order recording of the per-pair results (NA:758-1013),
 * check_for_two_interactions_on_same_edge (NA:528-632), the emission of the surviving basepairs (NA:1026-1044),
 * calculate_crossing_numbers (NA:1057-1179) and annotate_covalent_connections (NA:1181-1210).
 *
 * Output: one 16-byte row per tuple the reference appends to interaction_to_list_of_tuples, per structure in the
 * reference's own append order (so a stable grouping of a structure's rows by label gives the reference's dict:
 * same keys in the same order, same lists in the same order).  Labels are integer ids; the id space belongs to
 * the caller, who supplies the tables below (strings never reach the device). */
typedef struct {
    int32_t label;    /* label id; FR3D_LABEL_COVALENT | d for "p_<d>" */
    int32_t nt1;      /* global nt index (index_offset added) */
    int32_t nt2;
    int32_t crossing; /* crossing number, -1 = None (covalent rows) */
} fr3d_row_t;

#define FR3D_LABEL_COVALENT 0x40000000

/* Per label id. */
typedef struct {
    int32_t rev;   /* id of reverse_edges(label) (NA:446-465), -1 if not needed */
    int32_t near;  /* id of "n" + label (NA:1031-1032), -1 if not needed */
    uint32_t flags;/* bit0: rows are duplicated in reversed order (NA:1171-1177); bit1: label starts with "n";
                      bit2: label is one of cWW cWw cwW acWW acWw acwW (NA:1077); bits 8-15: the edge letter
                      label.replace("n","")[1].lower() (NA:543) */
    int32_t pad;
} fr3d_tail_label_t;

/* Per cutoff box (same indexing as the fr3d_box_t table). */
typedef struct {
    int32_t label[2][2]; /* [recorded for u2 ? 1 : 0][near]: the box name / "n" + name, and reverse_edges of it */
    uint64_t atoms1;     /* quality['atoms1'] / ['atoms2'] of store_basepair_quality (NA:2345-2361) as bit sets over */
    uint64_t atoms2;     /* the caller's numbering of the hydrogen-bond atom names */
} fr3d_tail_box_t;

typedef struct {
    int32_t n_labels;
    int32_t n_boxes;
    const fr3d_tail_label_t *labels;
    const fr3d_tail_box_t *boxes;
    const int32_t *slot_label; /* [8][32][2]: label ids of hit slot 0..7 and code: [0] the label recorded for
                                  (nt1, nt2) (for FR3D_SLOT_SO21 / SR21: for (nt2, nt1)), [1] the second label of
                                  the sO slots (interaction_reversed, NA:765,775), -1 elsewhere */
    uint64_t hydrogen_atoms;   /* bits of the atom names that contain "H" (NA:581-582) */
} fr3d_tail_tables_t;

/* Identity of the nts, as far as the tail reads it (index and parent come from nts->meta). */
typedef struct {
    int32_t n_struct;
    int32_t n_chains;
    int32_t n_ends;                  /* chain_ends_off[n_chains], known to the host (sizes the workspace) */
    int32_t max_struct_nt;           /* nts of the largest structure, known to the host; 0 = not announced.  From 1 024 nts the
                                        tail launches a second kernel with 1 024-thread blocks for the large structures */
    const int32_t *struct_off;       /* [n_struct + 1] first nt of each structure */
    const int32_t *struct_chain_off; /* [n_struct + 1] first chain-table entry of each structure */
    const int32_t *chain_ends_off;   /* [n_chains + 1] offset of the chain's nested-cWW endpoint array in the workspace;
                                        its length is the chain's largest index + 1 (NA:1106) */
    const int32_t *nt_chain;         /* [n_nt] chain-table entry = (structure, chain id): crossing numbers key chains by
                                        id only (NA:1068-1072) */
    const int32_t *nt_cov;           /* [n_nt] rank of the string model + " " + symmetry + " " + chain (NA:1192) */
} fr3d_tail_ident_t;

/* tail_counters (device int64[32]; [4..31] are written by profiling builds only): [0] total rows (may exceed row_cap: compare after the sync and re-run),
 * [1], [3] scratch, [2] != 0: a structure outside the supported ranges (more than 131 072 nts, an index outside
 * [0, 2^23), more than 65 536 chains or a label without the table entry it needs: FR3D_E_RANGE), [3] scratch.
 * n_hits is read from counters[1] on the device (clamped to hit_cap).  row_start / row_count [n_struct]: the rows of
 * structure s are rows[row_start[s] .. + row_count[s]) (the structures' blocks sit in no particular order). */
size_t fr3d_tail_workspace(int64_t hit_cap, int32_t n_nt, int32_t n_struct, int64_t n_ends);
int fr3d_annotation_tail(const fr3d_nts_t *nts, const fr3d_tail_ident_t *ident, const fr3d_tail_tables_t *tables,
                         const fr3d_hit_t *hits, int64_t hit_cap, const int64_t *counters, int32_t index_offset,
                         void *workspace, size_t workspace_bytes, fr3d_row_t *rows, int64_t row_cap,
                         int32_t *row_start, int32_t *row_count, int64_t *tail_counters, void *stream);

/* FR3D search screening.  Replaces fixed_radius_search (fr3d/search/pair_processing.py:106-177),
 * which get_pairlist (:7-104) calls once per search: every pair of points of the same model whose squared distance
 * is below max_cutoff^2.  xyz [n][3] (a NaN coordinate = no centre, the point is skipped, :137), model [n] (integer
 * ids of the reference's model strings).  Outputs, in arbitrary order: pair_a / pair_b = (pnt1, pnt2) in the
 * reference's orientation, pair_rank = first point of pnt1's bucket * 16 + stencil entry (sorting by (rank, pnt1,
 * pnt2) gives the reference's list order), dist2 = (x1-x2)^2 + (y1-y2)^2 + (z1-z2)^2 evaluated like the reference
 * (no fma).  counters (device int64[8]): [0] number of pairs (may exceed pair_cap: compare after the sync and re-run),
 * [2] != 0: a coordinate outside +-2046 cutoffs (FR3D_E_RANGE), [3] scratch. */
size_t fr3d_radius_search_workspace(int32_t n_points);
int fr3d_radius_search(const double *xyz, const int32_t *model, int32_t n_points, double max_cutoff, void *workspace,
                       size_t workspace_bytes, int32_t *pair_a, int32_t *pair_b, int32_t *pair_rank, double *dist2,
                       int64_t pair_cap, int64_t *counters, void *stream);

/* Glycosidic bond orientation.  Replaces the arithmetic of annotate_bond_orientation
 * (fr3d/classifiers/NA_unit_annotation.py:48-171) for every nt in one launch: chi = O4'-C1'-N9-C4 (purines) /
 * O4'-C1'-N1-C2 (pyrimidines) in degrees (NaN where an atom is missing or the residue has no parent) and
 * orientation = 0 NA, 1 anti, 2 syn, 3 int_syn (:150-155).  host_ring_slot[FR3D_N_PARENTS]: base slot of C4 / C2 per
 * parent (HOST pointer; N9 / N1 is slot 0).  Reads base_xyz, bb_xyz and the masks (for modified residues after
 * fr3d_frames_fit, which inserts the ideal-position copies the reference averages in). */
#define FR3D_ORIENT_NA 0
#define FR3D_ORIENT_ANTI 1
#define FR3D_ORIENT_SYN 2
#define FR3D_ORIENT_INT_SYN 3
int fr3d_bond_orientation(const fr3d_nts_t *nts, const int32_t *host_ring_slot, double *chi_degree, int32_t *orientation,
                          void *stream);

/* K5. Replaces besttransformation / besttransformation_weighted
 * (fr3d/geometry/superpositions.py:10-121) for n_prob independent problems.
 * Problem p uses rows off[p]..off[p+1]-1 of set1/set2/w (w may be NULL = unweighted).
 * Outputs (any may be NULL): U[p][9], mean1[p][3], mean2[p][3], rmsd[p], sse[p], new1[rows][3]. */
int fr3d_kabsch_batched(const double *set1, const double *set2, const double *w, const int32_t *off,
                        int32_t n_prob, double *U, double *mean1, double *mean2, double *rmsd,
                        double *sse, double *new1, void *stream);

/* K6. Replaces matrix_discrepancy (fr3d/geometry/discrepancy.py:165-233,
 * fr3d/search/discrepancy.py:165-233) over the double loop of fr3d/search/FR3D.py:433-442.
 * centers [N][n][3], rots [N][n][9], has_rot [N][n] (NULL = all present).
 * Writes out[(r - row0) * N + c] for row0 <= r < row1, 0 <= c < N; the diagonal is 0. */
int fr3d_discrepancy_all_vs_all(const double *centers, const double *rots, const uint8_t *has_rot,
                                int32_t N, int32_t n, double angle_weight, int32_t row0, int32_t row1,
                                double *out, void *stream);

/* The same for the multi-GPU split of the matrix: the upper triangle is dealt to the ranks in row blocks
 * (row0 a multiple of 32); a rank evaluates only the part of its block at or right of the diagonal,
 * out[(r - row0) * (N - row0) + (c - row0)] for row0 <= r < row1, row0 <= c < N - half the work of whole rows - and
 * fr3d_discrepancy_assemble places such a block into the N x N matrix and mirrors it (FR3D.py:439-441: the reference
 * fills both triangles from one evaluation too), on the rank that gathers the blocks. */
int fr3d_discrepancy_rows_upper(const double *centers, const double *rots, const uint8_t *has_rot, int32_t N,
                                int32_t n, double angle_weight, int32_t row0, int32_t row1, double *out, void *stream);
int fr3d_discrepancy_assemble(const double *block_upper, int32_t N, int32_t row0, int32_t row1, double *matrix,
                              void *stream);

/* Measurement aid (no reference counterpart): launches (SMs x 8) blocks of 256 threads that each run `iters` rounds of
 * 8 independent fp64 fused multiply-adds; *host_flops = the flops of the launch.  bench.py times it with CUDA events to
 * get the DFMA peak the discrepancy kernels' roofline is quoted against (MEASURED_PEAKS.json has no fp64 figure). */
int fr3d_fp64_fma_probe(int64_t iters, double *sink, int64_t *host_flops, void *stream);

/* Replaces matrix_discrepancy_cutoff (discrepancy.py:235-313) of one query against M candidates
 * (fr3d/search/search.py:1073-1085). out[m] = NaN where the reference returns None. */
int fr3d_discrepancy_one_vs_many(const double *q_centers, const double *q_rots, const uint8_t *q_has_rot,
                                 const double *centers, const double *rots, const uint8_t *has_rot,
                                 int32_t M, int32_t n, double cutoff /* < 0: no cutoff */, double angle_weight,
                                 const double *center_weight /* [n] or NULL */, double *out, void *stream);

/* ---- K8: ordering by similarity of the all-vs-all matrix ---------------------------------------
 * Replaces treePenalizedPathLength and its parts in both copies of the reference, fr3d/ordering/orderBySimilarity.py
 * (:10-41 treePenalty, :43-64 normalizePenaltyMatrix, :66-80 twoOptSwap, :82-115 greedyInsertionPathLength, :131-186,
 * :237-245) and fr3d/search/orderBySimilarity.py (:9-113; called at fr3d/search/FR3D.py:463), on a matrix resident in
 * HBM (e.g. the output of fr3d_discrepancy_all_vs_all).  All matrices are device double [n][n], row major; results are
 * bit-identical to the reference's floats (explicitly rounded operations in its operation order).  The starting orders
 * (random.shuffle of the GLOBAL Python generator) come from the caller. */
size_t fr3d_order_workspace(int32_t n, int32_t reps);   /* reps = 0: enough for fr3d_tree_penalty */
/* treePenalty: penalty[i][j] = height of the average-linkage merge (scipy linkage(squareform(d), "average"), restated as
 * its nearest-neighbour chain) that joins i and j.  flags (device int32[1]): 1 = not symmetric, 2 = non-zero diagonal,
 * 4 = non-finite entry - the inputs scipy refuses with ValueError; penalty is undefined then. */
int fr3d_tree_penalty(const double *distance, int32_t n, double *penalty, int32_t *flags, void *workspace,
                      size_t workspace_bytes, void *stream);
/* mode 0: out = distance + strength * penalty (search copy, strength 5); mode 1: penalty first scaled by
 * sum(distance) / sum(penalty) when that sum is positive (normalizePenaltyMatrix; the two sums run sequentially over the
 * rows like the reference's loops and are left in sums[0], sums[1], device double[2]); mode 2: out = the scaled penalty
 * alone. */
int fr3d_penalized_distance(const double *distance, const double *penalty, int32_t n, int32_t mode, double strength,
                            double *sums, double *out, void *stream);
/* One repetition per CTA.  mode bit 0: greedy insertion of start_orders[r][2..] into the path start_orders[r][0..1];
 * bit 1: twoOptSwap of the result (or of start_orders[r] itself when bit 0 is clear).  orders [reps][n], scores [reps]
 * (path score of the greedy insertion, 0 without it), reductions [reps] (the 2-opt's distance_reduction). */
int fr3d_order_paths(const double *distance, int32_t n, const int32_t *start_orders, int32_t reps, int32_t mode,
                     int32_t *orders, double *scores, double *reductions, void *workspace, size_t workspace_bytes,
                     void *stream);
/* reorderSymmetricMatrix: out[i][j] = out[j][i] = distance[order[i]][order[j]], j >= i (zero_diagonal: j > i, search copy). */
int fr3d_reorder_symmetric(const double *distance, int32_t n, const int32_t *order, int32_t zero_diagonal, double *out,
                           void *stream);

/* ---- K9: FR3D search, intersection of the pairwise constraint lists (VERDICT round 1, N3 second half) ------------------
 * Replaces getPossibilities / extendFragment (fr3d/search/search.py:263-375) over buildSecondElementList (:377-398): the
 * possibilities are grown one position per call, breadth first, one warp per fragment; output fragments are in
 * lexicographic order (the reference's order is that of Python set iteration; the SET is the same).
 * The query lives in DEVICE memory (the caller uploads this struct); [a][b] tables are indexed a * 16 + b, a < b. */
#define FR3D_SEARCH_MAX_POSITIONS 16
typedef struct {
    int32_t n_positions, n_units, symbolic /* Q["type"] == "symbolic" */, reserved;
    const int32_t *rowptr[FR3D_SEARCH_MAX_POSITIONS * FR3D_SEARCH_MAX_POSITIONS]; /* CSR of listOfPairs[a][b] by first element:
                                                  [n_units + 1]; NULL = "full" */
    const int32_t *cols[FR3D_SEARCH_MAX_POSITIONS * FR3D_SEARCH_MAX_POSITIONS];   /* second elements, ascending, unique */
    const uint8_t *universe[FR3D_SEARCH_MAX_POSITIONS];   /* [n_units] membership of universe[a]; needed where a list [a][b] is full */
    const double *centers;                                /* [n_units][3] units[u]["centers"]; geometric / mixed only */
    double perm_dist[FR3D_SEARCH_MAX_POSITIONS * FR3D_SEARCH_MAX_POSITIONS];      /* Q["permutedDistance"] */
    double cutoff[FR3D_SEARCH_MAX_POSITIONS];             /* Q["cutoff"][m] */
} fr3d_search_query_t;
/* The first pairs (search.py:351-366): pairs [n_pairs][2] = listOfPairs[0][1], sorted.  emit = 0: counts[i] (device
 * int64 [n_pairs + 1]) = exclusive prefix of "pair i starts a fragment", counts[n_pairs] = their number; emit = 1 (same
 * counts): out_frags [total][2], out_errs [total] (geometric / mixed: the pair's distance error; NULL for symbolic). */
int fr3d_possibilities_first(const fr3d_search_query_t *dev_query, const int32_t *pairs, int64_t n_pairs, int32_t emit,
                             int64_t *counts, int32_t *out_frags, double *out_errs, int64_t out_cap, void *stream);
/* One step of extendFragment for all fragments at once: frags [n_frags][level + 1] -> out_frags [total][level + 2] (position
 * level + 1 added), errs / out_errs the running distance errors.  emit = 0 counts (counts [n_frags + 1] as above), emit = 1
 * writes (nothing if counts[n_frags] > out_cap).  flags (device int32[1]) bit 0: the position has only "full" lists -
 * "Query is underspecified" (search.py:295-300). */
int fr3d_possibilities_extend(const fr3d_search_query_t *dev_query, int32_t level, const int32_t *frags, const double *errs,
                              int64_t n_frags, int32_t emit, int64_t *counts, int32_t *out_frags, double *out_errs,
                              int64_t out_cap, int32_t *flags, void *stream);

/* ---- native mmCIF ingest (host code) -------------------------------------------------------------
 * Replaces, for the nucleotides of a file, fr3d.cif.reader.Cif(handle).structure() (fr3d/cif/reader.py:440-629: operators,
 * one atom per (operator, atom_site row), residues by stable sort + group, alternate conformers) and the flattening of
 * the resulting Components into nt records (fr3d/data/components.py:118-202, base.py:150-166; previous O3' of
 * NA_pairwise_interactions.py:480-525) for many files at once, one thread per file.  Everything here is HOST memory. */
typedef struct fr3d_ingest fr3d_ingest_t;
const char *fr3d_ingest_last_error(void);
/* Slot tables, once per process: tab-separated lines "S seq parent flags dupmask", then "B atom slot" / "K atom slot" (base /
 * backbone map of that residue type), and "A parent heavy h1 h2" (amino-hydrogen slots); built by the Python layer. */
int fr3d_ingest_tables(const char *host_blob, size_t len);
/* Parse n_texts CIF texts with n_threads workers (0 = one per core).  *out must be released with fr3d_ingest_free. */
int fr3d_cif_parse(const char *const *host_texts, const int64_t *lengths, int32_t n_texts, int32_t n_threads,
                   fr3d_ingest_t **out);
/* sizes[0] = nts, [1] = structures, [2..8] = bytes of the string columns model, chain, sequence, symmetry, alt_id,
 * insertion_code (per nt) and pdb (per structure); values are NUL-terminated, "\1" stands for None. */
int fr3d_ingest_sizes(const fr3d_ingest_t *h, int64_t *sizes);
/* Fill caller-owned HOST buffers (pinned memory works): the fr3d_nts_t planes base_xyz [n][16][3], bb_xyz [n][10][3],
 * base_mask, bb_mask [n], meta [n][8]; struct_off [S + 1]; number, index [n] (index INT64_MIN = None); the seven string
 * columns (any may be NULL). */
int fr3d_ingest_fill(const fr3d_ingest_t *h, double *base_xyz, double *bb_xyz, uint32_t *base_mask, uint32_t *bb_mask,
                     int32_t *meta, int64_t *struct_off, int64_t *number, int64_t *index, char *const *string_columns);
void fr3d_ingest_free(fr3d_ingest_t *h);

#ifdef __cplusplus
}
#endif
#endif /* FR3D_B200_H */
