// K4 building blocks: helpers, the per-pair shared-memory scratch and the out-of-line rare paths of the
// one-warp-per-pair rule engine.  The kernel itself is in classify_pm.cuh (the first, pair-major kernel
// that lived here is in the git history; profiles/README.md has its ncu numbers).  Since the screened pipeline
// (classify_screen.cuh -> classify_tpp.cuh / classify_list2.cuh) that warp-per-pair kernel only serves callers
// that pass no workspace; the geometric helpers below (left_of, in_hull, in_ellipse, hb_angle...) serve every kernel.
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

// The same angle, but 0.0 (the reference's "no angle" value, which every caller skips) when it is certainly below
// 107.4 degrees: cos >= 0, or sin > 3.2 |cos| (tan 72.6 deg = 3.2).  Every base-phosphate / base-ribose label needs
// more than 110 degrees (NA:1947-1968), so the fp64 atan2 expansion only runs for the angles that can matter.
__device__ __forceinline__ double hb_angle_if_wide(const double *A, const double *B, const double *C) {
    const double u0 = A[0] - B[0], u1 = A[1] - B[1], u2 = A[2] - B[2];
    const double w0 = C[0] - B[0], w1 = C[1] - B[1], w2 = C[2] - B[2];
    const double cosang = u0 * w0 + u1 * w1 + u2 * w2;
    const double x0 = u1 * w2 - u2 * w1, x1 = u2 * w0 - u0 * w2, x2 = u0 * w1 - u1 * w0;
    const double sinang = sqrt(x0 * x0 + x1 * x1 + x2 * x2);
    if (cosang >= 0.0 || sinang > 3.2 * (-cosang)) return 0.0;
    return 180.0 * atan2_once(sinang, cosang) / 3.141592653589793;
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

// ---- modified residues carrying ideal-position copies of their base atoms (FR3D_FLAG_DUP) ----
// A "twin" slot stores the mean of the observed atom and its ideal copy (what centers[] returns); the
// reference's atoms() loops (calculate_basepair_gap NA:2211-2220, calculate_base_min_distances
// NA:1811-1827) see BOTH points.  ideal = centre + U . std, observed = 2 . mean - ideal.
__device__ __forceinline__ void ideal_point(const double *fr, int parent, int slot, double &i0, double &i1, double &i2) {
    const double s0 = c_tab.std_xyz[parent][slot][0], s1 = c_tab.std_xyz[parent][slot][1],
                 s2 = c_tab.std_xyz[parent][slot][2];
    i0 = fr[0] + (fr[3] * s0 + fr[4] * s1 + fr[5] * s2);
    i1 = fr[1] + (fr[6] * s0 + fr[7] * s1 + fr[8] * s2);
    i2 = fr[2] + (fr[9] * s0 + fr[10] * s1 + fr[11] * s2);
}

__device__ __noinline__ void gaps_twins(const WarpScratch &S, uint32_t bm1, uint32_t bm2, uint32_t tw1, uint32_t tw2,
                                        int par1, int par2, int lane, double &gap12, double &gap21) {
    const int half = lane >> 4, a = lane & 15;
    const double *own = half ? &S.fr[0] : &S.fr[12];      // frame of the nt the lane's atom belongs to
    const double *oth = half ? &S.fr[12] : &S.fr[0];      // frame it is seen from
    const bool valid = ((half ? bm1 : bm2) >> a) & 1u;
    const bool twin = valid && (((half ? tw1 : tw2) >> a) & 1u);
    const double p0 = S.atoms[lane][0], p1 = S.atoms[lane][1], p2 = S.atoms[lane][2];
    double i0 = 0, i1 = 0, i2 = 0;
    if (twin) ideal_point(own, half ? par1 : par2, a, i0, i1, i2);
    const double a0 = twin ? 2.0 * p0 - i0 : p0, a1 = twin ? 2.0 * p1 - i1 : p1, a2 = twin ? 2.0 * p2 - i2 : p2;
    double w0 = a0 - oth[0], w1 = a1 - oth[1], w2 = a2 - oth[2];
    double kA = valid ? (w0 * w0 + w1 * w1 + w2 * w2) : 1e300;
    const double zA = w0 * oth[5] + w1 * oth[8] + w2 * oth[11];
    w0 = i0 - oth[0]; w1 = i1 - oth[1]; w2 = i2 - oth[2];
    double kB = twin ? (w0 * w0 + w1 * w1 + w2 * w2) : 1e300;
    const double zB = w0 * oth[5] + w1 * oth[8] + w2 * oth[11];
    double g = 100.0;
#pragma unroll 1
    for (int r = 0; r < 3; ++r) {
        const bool useA = kA <= kB;
        double offer = useA ? kA : kB;
        const double offz = useA ? zA : zB;
        int idx = a;
        group_argmin<16>(offer, idx);
        const double zz = fabs(__shfl_sync(FULL, offz, (lane & 16) | idx));
        if (offer < 1e299) {
            if (zz < g) g = zz;
            if (a == idx) { if (useA) kA = 1e300; else kB = 1e300; }
        }
    }
    gap12 = __shfl_sync(FULL, g, 0);
    gap21 = __shfl_sync(FULL, g, 16);
}

__device__ __noinline__ double mind_twins(const WarpScratch &S, uint32_t bm1, uint32_t bm2, uint32_t tw1, uint32_t tw2,
                                          int par1, int par2, int lane) {
    const int half = lane >> 4, a = lane & 15;
    const bool v2 = (bm2 >> a) & 1u;
    const bool t2 = v2 && ((tw2 >> a) & 1u);
    const double q0 = S.atoms[a][0], q1 = S.atoms[a][1], q2 = S.atoms[a][2];
    double j0 = 0, j1 = 0, j2 = 0;
    if (t2) ideal_point(&S.fr[12], par2, a, j0, j1, j2);
    const double b0 = t2 ? 2.0 * q0 - j0 : q0, b1 = t2 ? 2.0 * q1 - j1 : q1, b2 = t2 ? 2.0 * q2 - j2 : q2;
    double best = 1e300;
#pragma unroll 1
    for (int s = 0; s < 8; ++s) {
        const int k1 = s + 8 * half;
        if (v2 && ((bm1 >> k1) & 1u)) {
            const bool t1 = (tw1 >> k1) & 1u;
            const double r0 = S.atoms[16 + k1][0], r1 = S.atoms[16 + k1][1], r2 = S.atoms[16 + k1][2];
            double i0 = 0, i1 = 0, i2 = 0;
            if (t1) ideal_point(&S.fr[0], par1, k1, i0, i1, i2);
            const double a0 = t1 ? 2.0 * r0 - i0 : r0, a1 = t1 ? 2.0 * r1 - i1 : r1, a2 = t1 ? 2.0 * r2 - i2 : r2;
            double d0 = a0 - b0, d1 = a1 - b1, d2 = a2 - b2;
            best = fmin(best, d0 * d0 + d1 * d1 + d2 * d2);
            if (t2) { d0 = a0 - j0; d1 = a1 - j1; d2 = a2 - j2; best = fmin(best, d0 * d0 + d1 * d1 + d2 * d2); }
            if (t1) {
                d0 = i0 - b0; d1 = i1 - b1; d2 = i2 - b2; best = fmin(best, d0 * d0 + d1 * d1 + d2 * d2);
                if (t2) { d0 = i0 - j0; d1 = i1 - j1; d2 = i2 - j2; best = fmin(best, d0 * d0 + d1 * d1 + d2 * d2); }
            }
        }
    }
    best = group_min<32>(best);
    return best < 1e299 ? sqrt(best) : 9999.0;
}

#undef C1
#undef U1
#undef C2
#undef U2

}  // namespace fr3d
