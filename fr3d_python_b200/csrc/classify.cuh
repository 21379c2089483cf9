// K4: one warp per candidate pair applies every pairwise rule of the reference
// annotator and emits compacted hit records.
//
// Replaces, per pair (nt1, nt2), in the reference's own evaluation order
// (NA = fr3d/classifiers/NA_pairwise_interactions.py):
//   check_base_oxygen_stack_rings x2   NA:1278-1473   (lanes 0-5 / 8-13: one oxygen each)
//   check_base_base_stacking           NA:1602-1739   (lanes 0-15: base atoms of nt2 in frame 1,
//                                                      lanes 16-31: base atoms of nt1 in frame 2)
//   check_sugar_ribose x2              NA:2262-2342   (lane 0 / lane 1)
//   check_base_backbone_interactions   NA:1867-2049   (lane = site*6 + oxygen; sticky selection on lane 0)
//   check_coplanar + calculate_basepair_gap + calculate_base_min_distances
//                                      NA:2051-2259, 1799-1829 (both half-warps at once)
//   check_basepair_cutoffs x2          NA:2364-2759   (one lane per cutoff box)
//   12-vs-21 arbitration               NA:935-957
// Everything that decides a threshold is fp64.  All branches are warp-uniform
// (a warp owns one pair), expensive stages (15x15 min distance, gaps) run only
// when the reference would reach them.
//
// Memory: per pair the warp stages both nts' frames (24 doubles), base atoms
// (32 x 3) and backbone atoms (20 x 3) in shared memory with coalesced loads
// (lane k reads element k's contiguous bytes); every rule then reads shared
// memory only.  The next pair's indices are prefetched.  Algorithmic traffic
// per pair (gather model): 2 x 760 B nt records + 12 B pair +
// 32 B hit slot.  Each rule family starts with a warp vote on a cheap
// NECESSARY condition (atom over the hull, oxygen inside the slab/disc that
// contains every ring, donor-oxygen distance below the near cutoff, frame-only
// coplanarity terms, |z| of the glycosidic displacement), so the expensive
// reductions run only for pairs that can still produce a label.
#pragma once

#include "tables.cuh"

namespace fr3d {

#define FULL 0xFFFFFFFFu

__device__ __forceinline__ bool left_of(const double *l, double x, double y) { return l[0] * x + l[1] * y + l[2] > 0; }

// check_convex_hull_atoms, NA:1475-1528
__device__ __forceinline__ bool in_hull(int parent, double x, double y, double z) {
    if (!(fabs(z) < 4.5)) return false;
    bool in = true;
#pragma unroll
    for (int k = 0; k < 7; ++k) in = in && left_of(c_tab.hull[parent][k], x, y);
    return in;
}

__device__ __forceinline__ bool in_ellipse(const double *e, double x, double y) {
    const double dx = x - e[2], dy = y - e[3];
    return e[0] * (dx * dx) + e[1] * dx * dy + dy * dy < e[4];
}

// min over a group of `width` lanes of (key, idx) pairs, lexicographic; all lanes get the result
template <int WIDTH>
__device__ __forceinline__ void group_argmin(double &key, int &idx) {
#pragma unroll
    for (int off = WIDTH / 2; off > 0; off >>= 1) {
        const double k2 = __shfl_xor_sync(FULL, key, off);
        const int i2 = __shfl_xor_sync(FULL, idx, off);
        if (k2 < key || (k2 == key && i2 < idx)) { key = k2; idx = i2; }
    }
}

template <int WIDTH>
__device__ __forceinline__ double group_min(double v) {
#pragma unroll
    for (int off = WIDTH / 2; off > 0; off >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, off));
    return v;
}

template <int WIDTH>
__device__ __forceinline__ double group_max(double v) {
#pragma unroll
    for (int off = WIDTH / 2; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, off));
    return v;
}

// The kernel is instruction-cache sensitive (ncu: stall_no_instruction dominated the first
// version, 9.9 k SASS instructions): the fp64 atan2 expansion and the rare, long scalar paths are
// kept out of line so that they exist once, and the orientation loop is not unrolled.
__device__ __noinline__ double atan2_once(double y, double x) { return atan2(y, x); }

// angle_between_vectors (NA:2873-2881) of (A-B, C-B), degrees
__device__ __forceinline__ double hb_angle(const double *A, const double *B, const double *C) {
    const double u0 = A[0] - B[0], u1 = A[1] - B[1], u2 = A[2] - B[2];
    const double w0 = C[0] - B[0], w1 = C[1] - B[1], w2 = C[2] - B[2];
    const double cosang = u0 * w0 + u1 * w1 + u2 * w2;
    const double x0 = u1 * w2 - u2 * w1, x1 = u2 * w0 - u0 * w2, x2 = u0 * w1 - u1 * w0;
    const double sinang = sqrt(x0 * x0 + x1 * x1 + x2 * x2);
    return 180.0 * atan2_once(sinang, cosang) / 3.141592653589793;
}

__device__ __forceinline__ double dist3(const double *a, const double *b) {
    const double d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
    return sqrt(d0 * d0 + d1 * d1 + d2 * d2);
}

static constexpr int WARPS_PER_BLOCK = 4;

// Per-warp staging area in shared memory: everything the rules read for one pair.
struct alignas(16) WarpScratch {
    double fr[24];        // c1[3] U1[9] c2[3] U2[9]
    double atoms[32][3];  // 0-15: base atoms of nt2, 16-31: base atoms of nt1   (offset 192, 16-byte aligned)
    double bb[20][3];     // 0-9: backbone atoms of nt2, 10-19: backbone atoms of nt1   (offset 960)
    double ang[32];
    double dst[32];
};

#define C1(k) (S.fr[(k)])
#define U1(k) (S.fr[3 + (k)])
#define C2(k) (S.fr[12 + (k)])
#define U2(k) (S.fr[15 + (k)])

// (p - c) . U for the frame stored at fr[0..11]   (translate_rotate_point, NA:1263-1275)
__device__ __forceinline__ void project(const double *fr, double px, double py, double pz, double &x, double &y,
                                        double &z) {
    const double v0 = px - fr[0], v1 = py - fr[1], v2 = pz - fr[2];
    x = v0 * fr[3] + v1 * fr[6] + v2 * fr[9];
    y = v0 * fr[4] + v1 * fr[7] + v2 * fr[10];
    z = v0 * fr[5] + v1 * fr[8] + v2 * fr[11];
}

// check_sugar_ribose, NA:2262-2342, after the shared O2'-O2' screen.  Evaluated by every lane on
// broadcast shared-memory reads (warp-uniform).  Returns -1 none, 0 cSR, 1 tSR.
// A / B: backbone blocks of nt_a / nt_b; base_pt: N3 / O2 of nt_a; (n0,n1,n2): base normal of nt_a.
__device__ __noinline__ int sugar_ribose_tail(const double (*A)[3], const double (*B)[3], uint32_t bba, uint32_t bbb,
                                              bool have_pt, const double *base_pt, double n0, double n1, double n2) {
    if (!have_pt) return -1;
    const double d0 = base_pt[0] - B[6][0], d1 = base_pt[1] - B[6][1], d2 = base_pt[2] - B[6][2];
    if (sqrt(d0 * d0 + d1 * d1 + d2 * d2) > 3.8) return -1;
    if (fabs(d0 * n0 + d1 * n1 + d2 * n2) > 2.0) return -1;
    // torsion C1'(a) C2'(a) C2'(b) C1'(b), NA:2904-2930
    if (!(bba >> 7 & 1u) || !(bba >> 8 & 1u) || !(bbb >> 7 & 1u) || !(bbb >> 8 & 1u)) return -1;
    double b0[3], b1[3], b2[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        b0[k] = -1.0 * (A[8][k] - A[7][k]);
        b1[k] = B[8][k] - A[8][k];
        b2[k] = B[7][k] - B[8][k];
    }
    const double nb1 = sqrt(b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2]);
    b1[0] /= nb1; b1[1] /= nb1; b1[2] /= nb1;
    const double d01 = b0[0] * b1[0] + b0[1] * b1[1] + b0[2] * b1[2];
    const double d21 = b2[0] * b1[0] + b2[1] * b1[1] + b2[2] * b1[2];
    double v[3], w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { v[k] = b0[k] - d01 * b1[k]; w[k] = b2[k] - d21 * b1[k]; }
    const double x = v[0] * w[0] + v[1] * w[1] + v[2] * w[2];
    const double cx0 = b1[1] * v[2] - b1[2] * v[1], cx1 = b1[2] * v[0] - b1[0] * v[2], cx2 = b1[0] * v[1] - b1[1] * v[0];
    const double y = cx0 * w[0] + cx1 * w[1] + cx2 * w[2];
    const double angle = atan2_once(y, x) * (180.0 / 3.141592653589793);
    return fabs(angle) > 90.0 ? 0 : 1;
}

// The sequential selection of check_base_backbone_interactions (NA:1925-2019) over the per-lane
// (angle, distance) values in S.ang / S.dst (index site*6 + oxygen; 0 = not a candidate).
// One lane runs it.  Returns bph | br << 8, each digit | near << 4, or 0xFF for none.
__device__ __noinline__ int backbone_select(const WarpScratch &S, int par1, uint32_t bb2) {
    int bph = -1, br = -1;
    // if nt2 has a P atom far from the plane of base 1, no BPh (NA:1915-1921)
    bool use_ph = true;
    if (bb2 & 1u) {
        const double *P = S.bb[0];
        if (P[0] != 0.0 || P[1] != 0.0 || P[2] != 0.0) {
            double tx, ty, tz;
            project(&S.fr[0], P[0], P[1], P[2], tx, ty, tz);
            if (fabs(tz) > 4.5) use_ph = false;
        }
    }
    const int nsites = c_tab.n_sites[par1];
    int n_true = 0, n_near = 0, n_keys = 0;
    unsigned long long site_ox = 0;        // 4 bits of oxygen indices per label digit
    double best_tq = -1e300, best_nq = -1e300;
    int best_tl = -1, best_nl = -1, first_tl = -1, first_nl = -1;
    int rt = 0, rn = 0, best_rtl = -1, best_rnl = -1, first_rtl = -1, first_rnl = -1;
    double best_rtq = -1e300, best_rnq = -1e300;
    double cutoff = 0.0, ncutoff = 0.0;
#pragma unroll 1
    for (int s = 0; s < nsites; ++s) {
        const int carbon = c_tab.site_carbon[par1][s];
        const int digit = c_tab.site_digit[par1][s];
        if (carbon == 1) { cutoff = 4.0; ncutoff = 4.5; }          // NA:1930-1935
        else if (carbon == 0) { cutoff = 3.5; ncutoff = 4.0; }
#pragma unroll 1
        for (int i = 0; i < 6; ++i) {
            const double an = S.ang[s * 6 + i], ds = S.dst[s * 6 + i];
            if (!(an != 0.0 && ds != 0.0)) continue;   // `if phosphateAngle:` / `if phosphateDistance:`
            if (i < 4) {                               // phosphate oxygens O5' OP1 OP2 prevO3' (NA:1938-1955)
                if (!use_ph) continue;
                if (an > 130.0 && ds < cutoff) {
                    const double q = (cutoff - ds) + (an - 130.0) / 20.0;
                    if (n_true == 0) first_tl = digit;
                    ++n_true;
                    if (q > best_tq || (q == best_tq && digit < best_tl)) { best_tq = q; best_tl = digit; }
                    if (!((site_ox >> (digit * 4)) & 0xFull)) ++n_keys;
                    site_ox |= 1ull << (digit * 4 + i);
                } else if (an > 110.0 && ds < ncutoff) {
                    const double q = (ncutoff - ds) + (an - 110.0) / 20.0;
                    if (n_near == 0) first_nl = digit;
                    ++n_near;
                    if (q > best_nq || (q == best_nq && digit < best_nl)) { best_nq = q; best_nl = digit; }
                }
            } else {                                   // ribose oxygens O2' O4' (NA:1958-1971)
                if (an > 130.0 && ds < cutoff) {
                    const double q = (cutoff - ds) + (an - 130.0) / 20.0;
                    if (rt == 0) first_rtl = digit;
                    ++rt;
                    if (q > best_rtq || (q == best_rtq && digit < best_rtl)) { best_rtq = q; best_rtl = digit; }
                } else if (an > 110.0 && ds < ncutoff) {
                    const double q = (ncutoff - ds) + (an - 110.0) / 20.0;
                    if (rn == 0) first_rnl = digit;
                    ++rn;
                    if (q > best_rnq || (q == best_rnq && digit < best_rnl)) { best_rnq = q; best_rnl = digit; }
                }
            }
        }
        // record the best phosphate interaction: sticky, evaluated after EVERY site (NA:1973-2003)
        if (n_true == 1) {
            bph = first_tl;
        } else if (n_keys > 1) {
            const unsigned o7 = (site_ox >> 28) & 0xF, o9 = (site_ox >> 36) & 0xF;
            const unsigned o3 = (site_ox >> 12) & 0xF, o5 = (site_ox >> 20) & 0xF;
            if (o7 && o9) { if (__popc(o7 | o9) > 1) bph = 8; }
            else if (o3 && o5) { if (__popc(o3 | o5) > 1) bph = 4; }
            if (bph < 0) bph = best_tl;
        } else if (n_true > 1) {
            bph = best_tl;
        } else if (n_near == 1) {
            bph = first_nl | 16;
        } else if (n_near > 1) {
            bph = best_nl | 16;
        }
        // best ribose interaction (NA:2005-2019)
        if (rt == 1) br = first_rtl;
        else if (rt > 1) br = best_rtl;
        else if (rn == 1) br = first_rnl | 16;
        else if (rn > 1) br = best_rnl | 16;
    }
    return (bph & 0xFF) | ((br & 0xFF) << 8);
}

#ifndef FR3D_CLASSIFY_MIN_BLOCKS
#define FR3D_CLASSIFY_MIN_BLOCKS 5   // 5 CTAs x 4 warps per SM: best of the {3,4,5,6} sweep (profiles/README.md)
#endif
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, FR3D_CLASSIFY_MIN_BLOCKS)
classify_pairs_kernel(fr3d_nts_t nts, const fr3d_box_t *__restrict__ boxes, const int32_t *__restrict__ box_index,
                      const int32_t *__restrict__ pair_i, const int32_t *__restrict__ pair_j,
                      const int32_t *__restrict__ pair_rank, int64_t pair_cap, uint32_t cats, int32_t index_offset,
                      fr3d_hit_t *__restrict__ hits, int64_t hit_cap, int64_t *counters) {
    __shared__ WarpScratch s_scratch[WARPS_PER_BLOCK];

    const int lane = threadIdx.x & 31;
    WarpScratch &S = s_scratch[threadIdx.x >> 5];
    const long long warp0 = (long long)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * WARPS_PER_BLOCK;
    long long n_pairs = counters[0];
    if (n_pairs > pair_cap) n_pairs = pair_cap;
    const int half = lane >> 4;      // 0: atoms of nt2 seen from frame 1, 1: atoms of nt1 seen from frame 2
    const int a = lane & 15;

    // software prefetch of the next pair's indices
    int nx1 = 0, nx2 = 0, nxr = 0;
    if (warp0 < n_pairs) { nx1 = pair_i[warp0]; nx2 = pair_j[warp0]; nxr = pair_rank[warp0]; }

#pragma unroll 1
    for (long long p = warp0; p < n_pairs; p += nwarps) {
        const int n1 = nx1, n2 = nx2, rank = nxr;
        if (p + nwarps < n_pairs) { nx1 = pair_i[p + nwarps]; nx2 = pair_j[p + nwarps]; nxr = pair_rank[p + nwarps]; }

        // ---- stage: frames (24 doubles), base atoms (32 x 3), backbone atoms (20 x 3), masks
        // absent slots hold zeros (packer / K1), so the atom loads do not wait for the masks
        const double *qa = nts.base_xyz + ((size_t)(half ? n1 : n2) * FR3D_BASE_SLOTS + a) * 3;
        const double px = qa[0], py = qa[1], pz = qa[2];
        double fv = 0.0;
        if (lane < 24) {
            const int n = lane < 12 ? n1 : n2;
            const int e = lane < 12 ? lane : lane - 12;
            fv = e < 3 ? nts.center[(size_t)n * 3 + e] : nts.rot[(size_t)n * 9 + (e - 3)];
        }
        double bx_ = 0, by_ = 0, bz_ = 0;
        if (lane < 20) {
            const double *q = nts.bb_xyz + ((size_t)(lane < 10 ? n2 : n1) * FR3D_BB_SLOTS + (lane < 10 ? lane : lane - 10)) * 3;
            bx_ = q[0]; by_ = q[1]; bz_ = q[2];
        }
        const uint32_t bm1 = nts.base_mask[n1], bm2 = nts.base_mask[n2];
        const uint32_t bb1 = nts.bb_mask[n1], bb2 = nts.bb_mask[n2];
        const int par1 = (int)((bm1 >> 16) & 15u) - 1, par2 = (int)((bm2 >> 16) & 15u) - 1;
        const bool valid = ((half ? bm1 : bm2) >> a) & 1u;
        __syncwarp();
        S.atoms[lane][0] = px; S.atoms[lane][1] = py; S.atoms[lane][2] = pz;
        if (lane < 24) S.fr[lane] = fv;
        if (lane < 20) { S.bb[lane][0] = bx_; S.bb[lane][1] = by_; S.bb[lane][2] = bz_; }
        __syncwarp();

        const double *fo = half ? &S.fr[12] : &S.fr[0];      // frame the lane's atom is projected into
        double x, y, z;
        project(fo, px, py, pz, x, y, z);
        const double nZ = U1(2) * U2(2) + U1(5) * U2(5) + U1(8) * U2(8);   // normal . normal = M[2][2] either way

        // decisions for the nine hit slots (warp-uniform): bit s of hmask, codes below
        unsigned hmask = 0;
        int c_so12 = 0, c_so21 = 0, c_stack = 0, c_sr12 = 0, c_sr21 = 0, c_bph = 0, c_br = 0, c_bp = 0;
        int bp_box = 0;
        double bp_cd = 0.0;

        // ================= base-oxygen stacking, both directions (NA:758-778) =================
        if (cats & FR3D_CAT_SO) {
            // group g (0: base nt1 / oxygens nt2, 1: base nt2 / oxygens nt1), o = lane&7 < 6
            const int g = (lane >> 3) & 1, o = lane & 7;
            const bool act = lane < 16 && o < 6;
            const int pb = g ? par2 : par1;
            // oxygens O2' O3' O4' O5' OP1 OP2 (NA:1290) -> backbone slots 6 5 4 3 1 2
            const int oslot = (0x213456 >> (4 * o)) & 7;
            const bool present = act && (((g ? bb1 : bb2) >> oslot) & 1u);
            double ox, oy, oz;
            const double *ob = S.bb[(g ? 10 : 0) + (act ? oslot : 0)];
            project(g ? &S.fr[12] : &S.fr[0], ob[0], ob[1], ob[2], ox, oy, oz);
            const bool has_tab = pb >= 0 && pb < FR3D_N_PARENTS;
            // every ring and near-ring ellipse lies inside the disc x^2 + y^2 < 25 (tests/test_abi.py checks
            // the tables), so an oxygen outside the slab |z| < 3.6 or the disc cannot score
            const bool maybe = present && has_tab && fabs(oz) < 3.6 && (ox * ox + oy * oy) < 25.0;
            if (__any_sync(FULL, maybe)) {
                bool ring = false, nr = false;
                if (maybe) {
                    const bool has_div = c_tab.has_div[pb];
                    const bool left = has_div && left_of(c_tab.ring_div[pb], ox, oy);
                    if (!(fabs(oz) < 2.0)) {             // impossibly close oxygens are skipped for rings only
                        ring = true;
                        if (left) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) ring = ring && left_of(c_tab.ring5[pb][k], ox, oy);
                        } else {
#pragma unroll
                            for (int k = 0; k < 6; ++k) ring = ring && left_of(c_tab.ring6[pb][k], ox, oy);
                        }
                    }
                    // near-ring ellipses, used only when no oxygen is over a ring (NA:1402-1459)
                    if (fabs(oz) < 3.5) nr = in_ellipse(left ? c_tab.ell5[pb] : c_tab.ell6[pb], ox, oy);
                }
                const unsigned ring_b = __ballot_sync(FULL, ring);
                const unsigned true_b = __ballot_sync(FULL, ring && fabs(oz) < 3.5);
                const unsigned gmask = 0xFFu << (g * 8);
                // over a ring: closest to the plane wins; true if any |z| < 3.5, else near (NA:1370-1400);
                // otherwise the ellipse hit closest to the base centre (NA:1440-1451)
                const bool use_ring = (ring_b & gmask) != 0;
                double key = use_ring ? (ring ? fabs(oz) : 1e300) : (nr ? (ox * ox + oy * oy) : 1e300);
                int idx = o;
                group_argmin<8>(key, idx);
                const double zsel = __shfl_sync(FULL, oz, (lane & ~7) | idx);
                int code = -1;
                if (key < 1e299) code = idx | ((zsel > 0) ? 0 : 8) | ((use_ring && (true_b & gmask)) ? 0 : 16);
                const int code12 = __shfl_sync(FULL, code, 0);
                const int code21 = __shfl_sync(FULL, code, 8);
                if (code12 >= 0) { hmask |= 1u << FR3D_SLOT_SO12; c_so12 = code12; }
                if (code21 >= 0) { hmask |= 1u << FR3D_SLOT_SO21; c_so21 = code21; }
            }
        }

        // ================= base-base stacking, part 1 (NA:1602-1702) =================
        int stack_code = -1;
        bool stack_both = false;
        if ((cats & FR3D_CAT_STACKING) && par1 >= 0 && par2 >= 0) {
            const bool inside = valid && in_hull(half ? par2 : par1, x, y, z);
            const unsigned in_b = __ballot_sync(FULL, inside);
            if (in_b) {                            // no atom over the other base: no overlap either way
                const double zmin = group_min<16>(valid ? z : 1e300);
                const double zmax = group_max<16>(valid ? z : -1e300);
                double key = valid ? fabs(z) : 1e300;
                int idx = a;
                group_argmin<16>(key, idx);
                const double zsel = __shfl_sync(FULL, z, (lane & 16) | idx);
                const unsigned hm = 0xFFFFu << (half * 16);
                const bool ov = (in_b & hm) && !(zmax > 0 && zmin < 0);       // return_overlap, NA:1567-1571
                const bool nt2on1 = __shfl_sync(FULL, (int)ov, 0);
                const bool nt1on2 = __shfl_sync(FULL, (int)ov, 16);
                const double z_2on1 = __shfl_sync(FULL, zsel, 0);
                const double z_1on2 = __shfl_sync(FULL, zsel, 16);
                if ((nt2on1 || nt1on2) && !(fabs(nZ) < 0.5)) {                 // NA:1717, cheap half first
                    const bool reverse = nt1on2 && !nt2on1;                    // NA:1663-1667
                    const double mvd = reverse ? z_1on2 : z_2on1;
                    int code = -1;                                             // 0 s35 1 s53 2 s33 3 s55
                    if (mvd > 0) { if (nZ > 0) code = 0; else if (nZ < 0) code = 2; }
                    else if (mvd < 0) { if (nZ > 0) code = 1; else if (nZ < 0) code = 3; }
                    if (reverse && code >= 0 && code < 2) code ^= 1;          // swap s35 <-> s53 (NA:1699-1702)
                    stack_code = code;
                    stack_both = nt2on1 && nt1on2;
                }
            }
        }

        // ================= sugar-ribose, both directions (NA:791-804) =================
        if ((cats & FR3D_CAT_SUGAR_RIBOSE) && par1 >= 0 && par2 >= 0 && par1 != 4 && par2 != 4 &&
            (bb1 >> 6 & 1u) && (bb2 >> 6 & 1u)) {
            // both directions start with the same O2'-O2' distance screen (NA:2281-2286)
            const double o0 = S.bb[16][0] - S.bb[6][0], o1 = S.bb[16][1] - S.bb[6][1], o2 = S.bb[16][2] - S.bb[6][2];
            const double oo = o0 * o0 + o1 * o1 + o2 * o2;
            if (oo < 14.4401 && !(sqrt(oo) > 3.8)) {
                const int s1 = c_tab.sr_slot[par1], s2 = c_tab.sr_slot[par2];
                const int r12 = sugar_ribose_tail(&S.bb[10], &S.bb[0], bb1, bb2, s1 >= 0 && ((bm1 >> s1) & 1u),
                                                  S.atoms[16 + (s1 < 0 ? 0 : s1)], U1(2), U1(5), U1(8));
                const int r21 = sugar_ribose_tail(&S.bb[0], &S.bb[10], bb2, bb1, s2 >= 0 && ((bm2 >> s2) & 1u),
                                                  S.atoms[s2 < 0 ? 0 : s2], U2(2), U2(5), U2(8));
                if (r12 >= 0) { hmask |= 1u << FR3D_SLOT_SR12; c_sr12 = r12; }
                if (r21 >= 0) { hmask |= 1u << FR3D_SLOT_SR21; c_sr21 = r21; }
            }
        }

        // ================= base-phosphate / base-ribose, nt1 base -> nt2 backbone (NA:807-829) =================
        if ((cats & FR3D_CAT_BACKBONE) && par1 >= 0) {
            const int site = lane / 6, ok = lane - site * 6;    // ok: 0 O5' 1 OP1 2 OP2 3 prevO3' | 4 O2' 5 O4'
            double dst = 0.0;
            bool cand = false;
            int hs = 0, ms = 0, oslot = 0;
            if (lane < 30 && site < c_tab.n_sites[par1]) {
                oslot = (0x469213 >> (4 * ok)) & 15;            // backbone slots 3 1 2 9 6 4
                hs = c_tab.site_h[par1][site];
                ms = c_tab.site_m[par1][site];
                if ((bb2 >> oslot & 1u) && (bm1 >> hs & 1u) && (bm1 >> ms & 1u)) {
                    const double *O = S.bb[oslot];
                    if (O[0] != 0.0 || O[1] != 0.0 || O[2] != 0.0) {            // .any(), NA:1939,1959
                        const double *Mv = S.atoms[16 + ms];
                        const double q0 = Mv[0] - O[0], q1 = Mv[1] - O[1], q2 = Mv[2] - O[2];
                        dst = q0 * q0 + q1 * q1 + q2 * q2;
                        // every true / near label needs distance < nCutoff (4.5 carbon, 4.0 nitrogen); the vote
                        // is on the square (a superset), the exact tests on the distance follow in the selection
                        cand = dst < (c_tab.site_carbon[par1][site] == 0 ? 16.0001 : 20.2501);
                    }
                }
            }
            if (__any_sync(FULL, cand)) {
                double ang = 0.0;
                if (cand) { ang = hb_angle(S.atoms[16 + ms], S.atoms[16 + hs], S.bb[oslot]); dst = sqrt(dst); }
                else dst = 0.0;                    // not a candidate: the selection skips it
                __syncwarp();
                S.ang[lane] = ang; S.dst[lane] = dst;
                __syncwarp();
                int sel = 0xFFFF;
                if (lane == 0) sel = backbone_select(S, par1, bb2);
                sel = __shfl_sync(FULL, sel, 0);
                if ((sel & 0xFF) != 0xFF) { hmask |= 1u << FR3D_SLOT_BPH; c_bph = sel & 0xFF; }
                if (((sel >> 8) & 0xFF) != 0xFF) { hmask |= 1u << FR3D_SLOT_BR; c_br = (sel >> 8) & 0xFF; }
            }
        }

        // ================= coplanar + basepair, part 1: frame-only terms and cutoff boxes =================
        // check_coplanar (NA:2051-2186) is a conjunction; its frame-only terms (:2098-2113) come first
        // here, the gap / min-distance ones are evaluated in part 2 only for pairs that pass them.
        // dot1 and dot2 swap roles between the two orientations and dot3 = |nZ| is symmetric.
        const bool gly_ok = (bm1 & 1u) && (bm2 & 1u) && par1 >= 0 && par2 >= 0;   // NA:674, 834-837
        bool cp_frames = false;
        unsigned ori_ok = 0;                     // bit ori: the parent combination is annotated (NA:842 / 894)
        unsigned bp_live = 0;                    // bit ori: some box is still within 1.0 after the angle term
        double cd0 = 1e300, cd1 = 1e300;         // per lane: running cutoff_distance of "my" box, per orientation
        if (gly_ok) {
            if (c_tab.combo_ok[par1 * FR3D_N_PARENTS + par2]) ori_ok |= 1u;
            if (c_tab.combo_ok[par2 * FR3D_N_PARENTS + par1]) ori_ok |= 2u;
            if ((cats & FR3D_CAT_COPLANAR) && ori_ok && fabs(nZ) > 0.7757) {
                // |e . n| / |e| < 0.3381 written without the square root and the three divisions
                const double e0 = C1(0) - C2(0), e1 = C1(1) - C2(1), e2 = C1(2) - C2(2);
                const double lim = (0.3381 * 0.3381) * (e0 * e0 + e1 * e1 + e2 * e2);
                const double d1 = e0 * U1(2) + e1 * U1(5) + e2 * U1(8);
                const double d2 = e0 * U2(2) + e1 * U2(5) + e2 * U2(8);
                cp_frames = d1 * d1 < lim && d2 * d2 < lim;
            }
            // glycosidic displacement: slot 0 = N9 / N1
            const double gd0 = S.atoms[0][0] - S.atoms[16][0], gd1 = S.atoms[0][1] - S.atoms[16][1],
                         gd2 = S.atoms[0][2] - S.atoms[16][2];
#pragma unroll 1
            for (int ori = 0; ori < 2; ++ori) {
                if (!((ori_ok >> ori) & 1u)) continue;
                const int combo = ori ? par2 * FR3D_N_PARENTS + par1 : par1 * FR3D_N_PARENTS + par2;
                const double *Ua = ori ? &S.fr[15] : &S.fr[3];
                const double *Ub = ori ? &S.fr[3] : &S.fr[15];
                const double s = ori ? -1.0 : 1.0;
                const double v0 = s * gd0, v1 = s * gd1, v2 = s * gd2;            // glycosidic_displacement
                const double dZ = v0 * Ua[2] + v1 * Ua[5] + v2 * Ua[8];            // displ12[2] (NA:847)
                if (fabs(dZ) > 3.6) continue;                                      // NA:2380
                const double dX = v0 * Ua[0] + v1 * Ua[3] + v2 * Ua[6];
                const double dY = v0 * Ua[1] + v1 * Ua[4] + v2 * Ua[7];
                const int sg = nZ > 0 ? 0 : 1;
                const int bstart = box_index[(combo * 2 + sg) * 2 + 0];
                const int bcount = box_index[(combo * 2 + sg) * 2 + 1];
                const bool mine = lane < bcount;
                const fr3d_box_t *bx = boxes + bstart + (mine ? lane : 0);
                double cd = 1e300;
                if (mine) {                                                        // NA:2546-2581
                    cd = 0.0;
                    cd += fmax(0.0, bx->xmin - dX);
                    cd += fmax(0.0, dX - bx->xmax);
                    cd += fmax(0.0, bx->ymin - dY);
                    cd += fmax(0.0, dY - bx->ymax);
                    if (bx->flags & 8) cd += fmax(0.0, sqrt(dX * dX + dY * dY) - bx->radiusmax);
                    cd += fmax(0.0, bx->zmin - dZ);
                    cd += fmax(0.0, dZ - bx->zmax);
                    cd += 3.0 * fmax(0.0, bx->normalmin - nZ);
                    cd += 3.0 * fmax(0.0, nZ - bx->normalmax);
                }
                bool ok = cd < 1.0;                                                // NA:2580
                if (!__any_sync(FULL, ok)) continue;                               // NA:2584
                const double M11 = Ua[1] * Ub[1] + Ua[4] * Ub[4] + Ua[7] * Ub[7];
                const double M01 = Ua[0] * Ub[1] + Ua[3] * Ub[4] + Ua[6] * Ub[7];
                double ang = atan2_once(M11, M01) * 57.29577951308232 - 90.0;      // NA:2594
                if (ang <= -90.0) ang += 360.0;
                if (ok) {
                    const double amin = bx->anglemin, amax = bx->anglemax;
                    if (amin < amax) {
                        cd += 0.05 * fmax(0.0, amin - ang);
                        cd += 0.05 * fmax(0.0, ang - amax);
                    } else {
                        cd += 0.05 * fmin(fmax(0.0, amin - ang), fmax(0.0, ang - amax));
                    }
                    ok = cd < 1.0;                                                 // NA:2611
                }
                if (!__any_sync(FULL, ok)) continue;                               // NA:2615
                bp_live |= 1u << ori;
                if (!ok) cd = 1e300;
                if (ori) cd1 = cd; else cd0 = cd;
            }
        }

        // ================= part 2: gaps and minimum base-base distance, once, only if needed =================
        double gap12 = 100.0, gap21 = 100.0, mind = 9999.0;
        const bool need_gaps = bp_live || cp_frames;
        if (need_gaps) {                          // calculate_basepair_gap, NA:2189-2259, both directions at once
            // squared distance to the other base centre orders the atoms like the reference's argsort
            double key = valid ? ((px - fo[0]) * (px - fo[0]) + (py - fo[1]) * (py - fo[1]) +
                                  (pz - fo[2]) * (pz - fo[2])) : 1e300;
            double g = 100.0;
#pragma unroll 1
            for (int r = 0; r < 3; ++r) {
                double k = key; int idx = a;
                group_argmin<16>(k, idx);
                const double zz = fabs(__shfl_sync(FULL, z, (lane & 16) | idx));
                if (k < 1e299) {
                    if (zz < g) g = zz;
                    if (a == idx) key = 1e300;
                }
            }
            gap12 = __shfl_sync(FULL, g, 0);
            gap21 = __shfl_sync(FULL, g, 16);
        }
        const bool cp_gap = cp_frames && (((ori_ok & 1u) && gap12 < 1.5179) || ((ori_ok & 2u) && gap21 < 1.5179));
        if (stack_code >= 0 || bp_live || cp_gap) {   // calculate_base_min_distances, NA:1799-1829
            // lower half: nt2 atom `a` against nt1 atoms 0-7; upper half: nt2 atom `a` against nt1 atoms 8-15
            const bool v2 = (bm2 >> a) & 1u;
            const double qx = S.atoms[a][0], qy = S.atoms[a][1], qz = S.atoms[a][2];
            double best = 1e300;
#pragma unroll 2
            for (int s = 0; s < 8; ++s) {
                const int k1 = s + 8 * half;
                if (v2 && ((bm1 >> k1) & 1u)) {
                    const double dx = S.atoms[16 + k1][0] - qx, dy = S.atoms[16 + k1][1] - qy,
                                 dz = S.atoms[16 + k1][2] - qz;
                    best = fmin(best, dx * dx + dy * dy + dz * dz);
                }
            }
            best = group_min<32>(best);
            mind = best < 1e299 ? sqrt(best) : 9999.0;
        }

        // ---- stacking, part 2 (NA:1707-1718)
        if (stack_code >= 0 && !(fabs(mind) < 1.0)) {
            int near = 1;
            if (fabs(mind) < 4.0 && fabs(mind) > 1.0 && fabs(nZ) > 0.6 && stack_both) near = 0;
            hmask |= 1u << FR3D_SLOT_STACK;
            c_stack = stack_code | (near << 2);
        }
        // ---- coplanar, part 2 (NA:2080-2096): orientation 12 or, failing that, 21 (NA:856-911)
        if (cp_gap) {
            const bool both_std = (bm1 >> 20 & 1u) && (bm2 >> 20 & 1u);
            if (mind < 3.4589 && !(mind >= 2.4589 && both_std)) hmask |= 1u << FR3D_SLOT_CP;
        }
        // ---- basepair, part 2: gap term, true / near classification, choice (NA:2642-2759)
        if (bp_live) {
            const double gmax = fmax(gap12, gap21);
            int box_sel[2] = {-1, -1}, near_sel[2] = {0, 0};
            double cd_sel[2] = {0.0, 0.0};
#pragma unroll
            for (int ori = 0; ori < 2; ++ori) {
                if (!((bp_live >> ori) & 1u)) continue;
                const int pa = ori ? par2 : par1, pb = ori ? par1 : par2;
                const int combo = pa * FR3D_N_PARENTS + pb;
                const int bstart = box_index[(combo * 2 + (nZ > 0 ? 0 : 1)) * 2 + 0];
                double cd = ori ? cd1 : cd0;
                const bool ok = cd < 1e299;
                const fr3d_box_t *bx = boxes + bstart + (ok ? lane : 0);
                bool is_match = false, is_near = false;
                int flags = 0;
                if (ok) {
                    flags = bx->flags;
                    const double gapmax = bx->gapmax;
                    if (gapmax > 0.1) cd += 4.0 * fmax(0.0, gmax - gapmax);        // NA:2643-2645
                    bool one_hb = false;                                           // NA:2648-2652
                    if ((flags & 2) && (pb == 1 || pb == 3)) one_hb = true;        // cSs with parent2 C or U
                    else if ((flags & 4) && pb == 3) one_hb = true;                // csS with parent2 U
                    const bool starts_n = flags & 1;
                    if (cd > 0) {
                        if (cd < 1.0 && !starts_n && (mind < 4.1 || one_hb)) is_near = true;
                    } else if (starts_n) {
                        is_near = true;
                    } else if (mind > 3.75 && !one_hb) {
                        is_near = true;
                    } else {
                        is_match = true;
                    }
                }
                const unsigned match_b = __ballot_sync(FULL, is_match);
                const unsigned near_b = __ballot_sync(FULL, is_near);
                if (match_b | near_b) {
                    // true matches: stable sort by (subcategory, len(name)) (NA:2682); else the nearest near
                    // match (NA:2685-2686); ties go to the earlier box like Python's stable sort
                    double key = 1e300;
                    if (match_b) { if (is_match) key = (double)(bx->subcat * 64 + bx->name_len); }
                    else if (is_near) key = cd;
                    int idx = lane;
                    group_argmin<32>(key, idx);
                    const double cds = __shfl_sync(FULL, cd, idx);
                    const int fls = __shfl_sync(FULL, flags, idx);
                    box_sel[ori] = bstart + idx;
                    cd_sel[ori] = cds;
                    // NA:2693-2696; bit1: an n-named box also carries "n" in its label
                    near_sel[ori] = ((cds > 0 && !(fls & 1)) ? 1 : 0) | ((fls & 1) ? 2 : 0);
                }
            }
            // ---- 12-vs-21 arbitration, NA:935-957
            int use = -1;
            if (box_sel[0] >= 0 && box_sel[1] >= 0) {
                const bool n0 = near_sel[0] != 0, n1_ = near_sel[1] != 0;
                // interaction12_reversed.lower() == interaction21.lower(): same "n" prefix state and same family
                const bool same = ((near_sel[0] & 1) == (near_sel[1] & 1)) &&
                                  (boxes[box_sel[0]].rev_lower_id == boxes[box_sel[1]].lower_id);
                if (same) use = 0;
                else if (n0 && n1_) use = (cd_sel[0] < cd_sel[1]) ? 0 : 1;
                else if (n1_) use = 0;
                else if (n0) use = 1;
                else use = 0;
            } else if (box_sel[0] >= 0) use = 0;
            else if (box_sel[1] >= 0) use = 1;
            if (use >= 0) {
                hmask |= 1u << FR3D_SLOT_BP;
                c_bp = (near_sel[use] & 1) | (use << 1);
                bp_box = box_sel[use];
                bp_cd = cd_sel[use];
            }
        }

        // ================= emit: lane s writes hit slot s, compacted =================
        if (hmask) {
            long long base = 0;
            if (lane == 0) base = (long long)atomicAdd((unsigned long long *)&counters[1], (unsigned long long)__popc(hmask));
            base = __shfl_sync(FULL, base, 0);
            if ((hmask >> lane) & 1u) {
                const long long pos = base + __popc(hmask & ((1u << lane) - 1u));
                if (pos < hit_cap) {
                    int mycode = 0;
                    switch (lane) {
                        case FR3D_SLOT_SO12: mycode = c_so12; break;
                        case FR3D_SLOT_SO21: mycode = c_so21; break;
                        case FR3D_SLOT_STACK: mycode = c_stack; break;
                        case FR3D_SLOT_SR12: mycode = c_sr12; break;
                        case FR3D_SLOT_SR21: mycode = c_sr21; break;
                        case FR3D_SLOT_BPH: mycode = c_bph; break;
                        case FR3D_SLOT_BR: mycode = c_br; break;
                        case FR3D_SLOT_BP: mycode = c_bp; break;
                        default: break;
                    }
                    const bool isbp = lane == FR3D_SLOT_BP;
                    const double cdv = isbp ? bp_cd : 0.0, gv = isbp ? fmax(gap12, gap21) : 0.0;
                    const uint32_t w3 = (uint32_t)lane | ((uint32_t)mycode << 8) | ((uint32_t)(isbp ? bp_box : 0) << 16);
                    uint4 *dst = reinterpret_cast<uint4 *>(hits + pos);
                    dst[0] = make_uint4((uint32_t)(n1 + index_offset), (uint32_t)(n2 + index_offset),
                                        (uint32_t)(rank + 16 * index_offset), w3);
                    dst[1] = make_uint4((uint32_t)__double2loint(cdv), (uint32_t)__double2hiint(cdv),
                                        (uint32_t)__double2loint(gv), (uint32_t)__double2hiint(gv));
                }
            }
        }
    }
}

#undef C1
#undef U1
#undef C2
#undef U2

}  // namespace fr3d
