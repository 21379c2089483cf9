// K4, thread-per-pair form.
//
// ncu on the warp-per-pair kernels (classify_pm.cuh, profiles/README.md) shows that the rule engine is
// bound by instruction issue, and that most of the ~950 warp instructions spent per pair are
// warp-UNIFORM bookkeeping (address arithmetic, mask decoding, votes, loop control, the scalar
// coplanar / basepair-box / sugar-ribose logic) executed by 32 lanes for one pair.  Here one THREAD owns
// one candidate pair and walks the same rules sequentially, so that bookkeeping is amortised over 32
// pairs per warp instruction, while the data-parallel parts (32 atom projections, 12 oxygens, 30
// donor/oxygen combinations, 256 atom-atom distances) become unrolled loops with plenty of
// instruction-level parallelism.
//
// Data access: this kernel reads the nt records straight from global memory, which is fine for the two short
// lists it serves (sO / sugar-ribose and base-phosphate) and for the single-launch A/B mode.  For the big lists
// that pattern is L1-wavefront bound (every load instruction touches 32 cache lines: ~80 % of the L1 wavefront
// rate with the issue slots 90 % idle), so they go through classify_list2.cuh, which stages the records of a
// warp's 32 pairs in shared memory with coalesced cp.async copies.  The helpers below take plain pointers and
// serve both kernels.
//
// The arithmetic is the same as in the warp-per-pair kernels and as in the reference; every function
// below cites the reference lines it restates (NA = fr3d/classifiers/NA_pairwise_interactions.py).
#pragma once

#include "classify.cuh"
#include "classify_pm.cuh"

namespace fr3d {

struct TFrame {
    double c0, c1, c2;
    double u0, u1, u2, u3, u4, u5, u6, u7, u8;     // row-major rotation_matrix
};

// frame from a staged row: centre 0..2, rotation 3..11
__device__ __forceinline__ TFrame load_tframe_row(const double *R) {
    TFrame f;
    f.c0 = R[0]; f.c1 = R[1]; f.c2 = R[2];
    f.u0 = R[3]; f.u1 = R[4]; f.u2 = R[5]; f.u3 = R[6]; f.u4 = R[7]; f.u5 = R[8]; f.u6 = R[9]; f.u7 = R[10]; f.u8 = R[11];
    return f;
}

__device__ __forceinline__ TFrame load_tframe(const fr3d_nts_t &nts, int n) {
    const double *c = nts.center + (size_t)n * 3;
    const double *r = nts.rot + (size_t)n * 9;
    TFrame f;
    f.c0 = c[0]; f.c1 = c[1]; f.c2 = c[2];
    f.u0 = r[0]; f.u1 = r[1]; f.u2 = r[2]; f.u3 = r[3]; f.u4 = r[4]; f.u5 = r[5]; f.u6 = r[6]; f.u7 = r[7]; f.u8 = r[8];
    return f;
}

// translate_rotate_point (NA:1263-1275): (p - centre) . U as a row vector
__device__ __forceinline__ void tproject(const TFrame &f, double px, double py, double pz, double &x, double &y,
                                         double &z) {
    const double v0 = px - f.c0, v1 = py - f.c1, v2 = pz - f.c2;
    x = v0 * f.u0 + v1 * f.u3 + v2 * f.u6;
    y = v0 * f.u1 + v1 * f.u4 + v2 * f.u7;
    z = v0 * f.u2 + v1 * f.u5 + v2 * f.u8;
}

// return_overlap (NA:1530-1571): base atoms of nt_b seen from the frame of nt_a.  hull = the 7 half planes of
// nt_a's parent ([7][3]).  Pass 1 is branch free over all 16 slots: the height z of every atom (min / max /
// closest to the plane) and a bit for the atoms inside the cylinder |z| < 4.5, x^2 + y^2 < 16 that contains every
// hull (tests/test_abi.py).  "inside" is an OR over the atoms, so pass 2 runs the exact half-plane test on the
// cylinder atoms only until the first one is in - a thread-per-pair warp pays for a branch if ANY lane takes it,
// and on the stacking list that is one or two iterations instead of sixteen.  An atom inside the disc of squared
// radius r2in that the hull contains (c_tab.hull_r2in, tests/test_abi.py) is in without any half-plane test.
__device__ __forceinline__ bool t_overlap(const TFrame &fa, const double *hull, double r2in, const double *q,
                                          uint32_t bmb, double &zsel) {
    double zmin = 1e300, zmax = -1e300, best = 1e300;
    unsigned cand = 0;
    bool inside = false;
    zsel = 0.0;
#pragma unroll 4
    for (int a = 0; a < FR3D_BASE_SLOTS; ++a) {
        const bool have = (bmb >> a) & 1u;
        const double v0 = q[a * 3] - fa.c0, v1 = q[a * 3 + 1] - fa.c1, v2 = q[a * 3 + 2] - fa.c2;
        const double z = v0 * fa.u2 + v1 * fa.u5 + v2 * fa.u8;                // the z of translate_rotate_point
        const double r2 = (v0 * v0 + v1 * v1 + v2 * v2) - z * z;              // x^2 + y^2 (U is orthonormal)
        if (have && fabs(z) < 4.5 && r2 < 16.0 + 1e-6) cand |= 1u << a;
        if (have && fabs(z) < 4.5 && r2 < r2in) inside = true;
        if (have && fabs(z) < best) { best = fabs(z); zsel = z; }
        if (have) { zmin = fmin(zmin, z); zmax = fmax(zmax, z); }
    }
    while (cand && !inside) {
        const int a = __ffs(cand) - 1;
        cand &= cand - 1;
        double x, y, z;
        tproject(fa, q[a * 3], q[a * 3 + 1], q[a * 3 + 2], x, y, z);
        bool in = fabs(z) < 4.5;
#pragma unroll
        for (int k = 0; k < 7; ++k) in = in && left_of(hull + k * 3, x, y);
        inside = in;
    }
    return inside && !(zmax > 0 && zmin < 0);
}

// check_base_oxygen_stack_rings (NA:1278-1473): base of nt_a, oxygens of nt_b.  Returns the code or -1.
__device__ __forceinline__ int t_so(const TFrame &fa, int pa, const double *B, uint32_t bbb) {
    if (pa < 0 || pa >= FR3D_N_PARENTS) return -1;
    const bool has_div = c_tab.has_div[pa];
    bool true_found = false, ring_found = false;
    double zmin_abs = 1e300, zring = 0.0;
    int oring = -1;
    double r2min = 1e300, znear = 0.0;
    int onear = -1;
#pragma unroll
    for (int o = 0; o < 6; ++o) {
        const int oslot = (0x213456 >> (4 * o)) & 7;          // O2' O3' O4' O5' OP1 OP2 -> slots 6 5 4 3 1 2
        if (!((bbb >> oslot) & 1u)) continue;
        double x, y, z;
        tproject(fa, B[oslot * 3], B[oslot * 3 + 1], B[oslot * 3 + 2], x, y, z);
        // every ring and near-ring ellipse lies inside the disc x^2 + y^2 < 25 (tests/test_abi.py)
        if (!(fabs(z) < 3.6) || !(x * x + y * y < 25.0)) continue;
        const bool left = has_div && left_of(c_tab.ring_div[pa], x, y);
        if (!(fabs(z) < 2.0)) {
            bool ring = true;
            if (left) {
#pragma unroll
                for (int k = 0; k < 4; ++k) ring = ring && left_of(c_tab.ring5[pa][k], x, y);
            } else {
#pragma unroll
                for (int k = 0; k < 6; ++k) ring = ring && left_of(c_tab.ring6[pa][k], x, y);
            }
            if (ring) {
                ring_found = true;
                if (fabs(z) < 3.5) true_found = true;
                if (fabs(z) < zmin_abs) { zmin_abs = fabs(z); zring = z; oring = o; }
            }
        }
        if (fabs(z) < 3.5 && in_ellipse(left ? c_tab.ell5[pa] : c_tab.ell6[pa], x, y)) {
            const double r2 = x * x + y * y;
            if (r2 < r2min) { r2min = r2; znear = z; onear = o; }
        }
    }
    if (ring_found) return oring | ((zring > 0) ? 0 : 8) | (true_found ? 0 : 16);
    if (onear >= 0) return onear | ((znear > 0) ? 0 : 8) | 16;
    return -1;
}

// check_sugar_ribose (NA:2262-2342) after the shared O2'-O2' screen.  -1 none, 0 cSR, 1 tSR.
__device__ __noinline__ int t_sr_tail(const double *A, const double *B, const double *base_a, uint32_t bma,
                                      uint32_t bba, uint32_t bbb, int pa, double n0, double n1, double n2) {
    const int bs = c_tab.sr_slot[pa];
    if (bs < 0 || !((bma >> bs) & 1u)) return -1;
    const double *bp = base_a + bs * 3;
    const double d0 = bp[0] - B[18], d1 = bp[1] - B[19], d2 = bp[2] - B[20];
    if (sqrt(d0 * d0 + d1 * d1 + d2 * d2) > 3.8) return -1;
    if (fabs(d0 * n0 + d1 * n1 + d2 * n2) > 2.0) return -1;
    if (!(bba >> 7 & 1u) || !(bba >> 8 & 1u) || !(bbb >> 7 & 1u) || !(bbb >> 8 & 1u)) return -1;
    double b0[3], b1[3], b2[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        b0[k] = -1.0 * (A[24 + k] - A[21 + k]);
        b1[k] = B[24 + k] - A[24 + k];
        b2[k] = B[21 + k] - B[24 + k];
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

// check_base_backbone_interactions (NA:1867-2049): nt1 base -> nt2 backbone, with the reference's own
// sequential, sticky selection.  Returns bph | br << 8 (digit | near << 4; 0xFF none).
// combos: bit site * 6 + i set = the (site, oxygen i) combination can yield a label (all ones: test every one).  The
// single-precision screen (classify_screen32.cuh) hands over the combinations that pass its necessary conditions
// (both site atoms and the oxygen present, massive atom within 4.5 A + margin, angle not certainly below 107 degrees),
// so only those are evaluated here.  A site without such a combination changes none of the selection's inputs, and
// the selection is idempotent (it only assigns values its inputs determine, or keeps a value already set), so
// skipping it for those sites gives the reference's result.
__device__ __noinline__ int t_backbone(const TFrame &f1, int par1, const double *A, const double *B, uint32_t bm1,
                                       uint32_t bb2, unsigned combos = 0xFFFFFFFFu) {
    int bph = -1, br = -1;
    bool use_ph = true;                    // P far from the plane of base 1: no BPh (NA:1915-1921)
    if (bb2 & 1u) {
        if (B[0] != 0.0 || B[1] != 0.0 || B[2] != 0.0) {
            double tx, ty, tz;
            tproject(f1, B[0], B[1], B[2], tx, ty, tz);
            if (fabs(tz) > 4.5) use_ph = false;
        }
    }
    const int nsites = c_tab.n_sites[par1];
    int n_true = 0, n_near = 0, n_keys = 0;
    unsigned long long site_ox = 0;
    double best_tq = -1e300, best_nq = -1e300;
    int best_tl = -1, best_nl = -1, first_tl = -1, first_nl = -1;
    int rt = 0, rn = 0, best_rtl = -1, best_rnl = -1, first_rtl = -1, first_rnl = -1;
    double best_rtq = -1e300, best_rnq = -1e300;
    double cutoff = 0.0, ncutoff = 0.0;
#pragma unroll 1
    for (int s = 0; s < nsites; ++s) {
        const int carbon = c_tab.site_carbon[par1][s];
        const int digit = c_tab.site_digit[par1][s];
        if (carbon == 1) { cutoff = 4.0; ncutoff = 4.5; }          // NA:1930-1935 (sticky: a site with neither keeps the last)
        else if (carbon == 0) { cutoff = 3.5; ncutoff = 4.0; }
        const unsigned site_combos = (combos >> (s * 6)) & 63u;
        if (!site_combos) continue;
        const int hs = c_tab.site_h[par1][s], ms = c_tab.site_m[par1][s];
        if ((bm1 >> hs & 1u) && (bm1 >> ms & 1u)) {
            const double *H = A + hs * 3, *Mv = A + ms * 3;
#pragma unroll 1
            for (int i = 0; i < 6; ++i) {
                if (!((site_combos >> i) & 1u)) continue;
                if (i < 4 && !use_ph) continue;
                const int oslot = (0x469213 >> (4 * i)) & 15;       // O5' OP1 OP2 prevO3' | O2' O4'
                if (!((bb2 >> oslot) & 1u)) continue;
                const double *O = B + oslot * 3;
                if (!(O[0] != 0.0 || O[1] != 0.0 || O[2] != 0.0)) continue;      // .any(), NA:1939,1959
                const double q0 = Mv[0] - O[0], q1 = Mv[1] - O[1], q2 = Mv[2] - O[2];
                const double d2 = q0 * q0 + q1 * q1 + q2 * q2;
                if (!(d2 < 20.2501)) continue;                     // every label needs distance < 4.5
                const double ds = sqrt(d2);
                const double an = hb_angle_if_wide(Mv, H, O);
                if (!(an != 0.0 && ds != 0.0)) continue;
                if (i < 4) {
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
                } else {
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
        }
        // sticky selection, evaluated after EVERY site (NA:1973-2019)
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
        if (rt == 1) br = first_rtl;
        else if (rt > 1) br = best_rtl;
        else if (rn == 1) br = first_rnl | 16;
        else if (rn > 1) br = best_rnl | 16;
    }
    return (bph & 0xFF) | ((br & 0xFF) << 8);
}

// cheap necessary condition for t_backbone: some donor heavy atom of nt1 within 4.5 A of a backbone oxygen
__device__ __forceinline__ bool t_backbone_possible(int par1, const double *A, const double *B, uint32_t bm1,
                                                    uint32_t bb2) {
    const int nsites = c_tab.n_sites[par1];
    bool any = false;
#pragma unroll 1
    for (int s = 0; s < nsites; ++s) {
        const int ms = c_tab.site_m[par1][s];
        if (!((bm1 >> ms) & 1u)) continue;
        const double m0 = A[ms * 3], m1 = A[ms * 3 + 1], m2 = A[ms * 3 + 2];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int oslot = (0x469213 >> (4 * i)) & 15;
            const double q0 = m0 - B[oslot * 3], q1 = m1 - B[oslot * 3 + 1], q2 = m2 - B[oslot * 3 + 2];
            if (((bb2 >> oslot) & 1u) && (q0 * q0 + q1 * q1 + q2 * q2 < 20.2501)) any = true;
        }
    }
    return any;
}

// one point of a (possibly twin) slot: v < 16 -> the observed / stored point, v >= 16 -> the ideal copy
__device__ __forceinline__ bool t_point(const TFrame &f, int par, const double *base, uint32_t bm, uint32_t tw,
                                        int v, double &p0, double &p1, double &p2) {
    const int k = v & 15;
    if (!((bm >> k) & 1u)) return false;
    const bool twin = (tw >> k) & 1u;
    if (v >= 16 && !twin) return false;
    const double *q = base + k * 3;
    p0 = q[0]; p1 = q[1]; p2 = q[2];
    if (twin) {
        const double s0 = c_tab.std_xyz[par][k][0], s1 = c_tab.std_xyz[par][k][1], s2 = c_tab.std_xyz[par][k][2];
        const double i0 = f.c0 + (f.u0 * s0 + f.u1 * s1 + f.u2 * s2);
        const double i1 = f.c1 + (f.u3 * s0 + f.u4 * s1 + f.u5 * s2);
        const double i2 = f.c2 + (f.u6 * s0 + f.u7 * s1 + f.u8 * s2);
        if (v >= 16) { p0 = i0; p1 = i1; p2 = i2; }
        else { p0 = 2.0 * p0 - i0; p1 = 2.0 * p1 - i1; p2 = 2.0 * p2 - i2; }
    }
    return true;
}

// calculate_basepair_gap (NA:2189-2259): min |z| (frame of nt_a) over the 3 base atoms of nt_b nearest
// to the centre of nt_a
__device__ __forceinline__ double t_gap_inline(const TFrame &fa, const TFrame &fb, int par_b, const double *base_b,
                                               uint32_t bmb, uint32_t twb) {
    double k0 = 1e300, k1 = 1e300, k2 = 1e300, z0 = 100.0, z1 = 100.0, z2 = 100.0;
    if (!twb) {
        // no ideal-copy twins (the usual case): branch-free insertion into the three smallest keys; an absent atom
        // gets key 1e300, which is never below the initial keys, so it changes nothing (the reference skips it)
#pragma unroll 4
        for (int v = 0; v < FR3D_BASE_SLOTS; ++v) {
            const double w0 = base_b[v * 3] - fa.c0, w1 = base_b[v * 3 + 1] - fa.c1, w2 = base_b[v * 3 + 2] - fa.c2;
            const double key = ((bmb >> v) & 1u) ? w0 * w0 + w1 * w1 + w2 * w2 : 1e300;
            const double z = fabs(w0 * fa.u2 + w1 * fa.u5 + w2 * fa.u8);
            const bool lt0 = key < k0, lt1 = key < k1, lt2 = key < k2;
            k2 = lt1 ? k1 : (lt2 ? key : k2);
            z2 = lt1 ? z1 : (lt2 ? z : z2);
            k1 = lt0 ? k0 : (lt1 ? key : k1);
            z1 = lt0 ? z0 : (lt1 ? z : z1);
            k0 = lt0 ? key : k0;
            z0 = lt0 ? z : z0;
        }
        return fmin(z0, fmin(z1, z2));
    }
    const int nv = twb ? 32 : 16;
#pragma unroll 1
    for (int v = 0; v < nv; ++v) {
        double p0, p1, p2;
        if (!t_point(fb, par_b, base_b, bmb, twb, v, p0, p1, p2)) continue;
        const double w0 = p0 - fa.c0, w1 = p1 - fa.c1, w2 = p2 - fa.c2;
        const double key = w0 * w0 + w1 * w1 + w2 * w2;
        const double z = fabs(w0 * fa.u2 + w1 * fa.u5 + w2 * fa.u8);
        if (key < k0) { k2 = k1; z2 = z1; k1 = k0; z1 = z0; k0 = key; z0 = z; }
        else if (key < k1) { k2 = k1; z2 = z1; k1 = key; z1 = z; }
        else if (key < k2) { k2 = key; z2 = z; }
    }
    return fmin(z0, fmin(z1, z2));
}

// out-of-line form (keeps the generic kernel's code size down)
__device__ __noinline__ double t_gap(const TFrame &fa, const TFrame &fb, int par_b, const double *base_b,
                                     uint32_t bmb, uint32_t twb) {
    return t_gap_inline(fa, fb, par_b, base_b, bmb, twb);
}

// calculate_base_min_distances (NA:1799-1829), general form (twin slots of modified residues included); the
// register-resident form for the big lists is t_mind_load / t_mind_scan in classify_list2.cuh
__device__ __forceinline__ double t_mind(const TFrame &f1, const TFrame &f2, int par1, int par2, const double *base1,
                                         const double *base2, uint32_t bm1, uint32_t bm2, uint32_t tw1, uint32_t tw2) {
    double best = 1e300;
    const int nv1 = tw1 ? 32 : 16, nv2 = tw2 ? 32 : 16;
#pragma unroll 1
    for (int v = 0; v < nv1; ++v) {
        double a0, a1, a2;
        if (!t_point(f1, par1, base1, bm1, tw1, v, a0, a1, a2)) continue;
        if (!tw2) {
            const double *q = base2;
#pragma unroll 4
            for (int k = 0; k < FR3D_BASE_SLOTS; ++k) {
                const double d0 = q[k * 3] - a0, d1 = q[k * 3 + 1] - a1, d2 = q[k * 3 + 2] - a2;
                const double dd = d0 * d0 + d1 * d1 + d2 * d2;
                if (((bm2 >> k) & 1u) && dd < best) best = dd;
            }
        } else {
#pragma unroll 1
            for (int w = 0; w < nv2; ++w) {
                double b0, b1, b2;
                if (!t_point(f2, par2, base2, bm2, tw2, w, b0, b1, b2)) continue;
                const double d0 = b0 - a0, d1 = b1 - a1, d2 = b2 - a2;
                best = fmin(best, d0 * d0 + d1 * d1 + d2 * d2);
            }
        }
    }
    return best < 1e299 ? sqrt(best) : 9999.0;
}

// check_basepair_cutoffs (NA:2546-2612) over the boxes of one orientation, in two passes.
struct TBp {
    int box;        // chosen box or -1
    int near;       // bit0 "n" prefix, bit1 n-named box
    double cd;
};

// L1 distance of one box up to and including the normal term (NA:2546-2581), in the reference's summation order
// (every term is >= 0, so once the x / y terms reach 1.0 the box is out whatever follows: return early)
__device__ __forceinline__ double t_box_cd(const fr3d_box_t *bx, double dX, double dY, double dZ, double nZ) {
    double cd = 0.0;
    cd += fmax(0.0, bx->xmin - dX);
    cd += fmax(0.0, dX - bx->xmax);
    cd += fmax(0.0, bx->ymin - dY);
    cd += fmax(0.0, dY - bx->ymax);
    if (!(cd < 1.0)) return cd;
    if (bx->flags & 8) cd += fmax(0.0, sqrt(dX * dX + dY * dY) - bx->radiusmax);
    cd += fmax(0.0, bx->zmin - dZ);
    cd += fmax(0.0, dZ - bx->zmax);
    cd += 3.0 * fmax(0.0, bx->normalmin - nZ);
    cd += 3.0 * fmax(0.0, nZ - bx->normalmax);
    return cd;
}

// ... plus the angle term (NA:2594-2611)
__device__ __forceinline__ double t_box_angle(const fr3d_box_t *bx, double cd, double ang) {
    const double amin = bx->anglemin, amax = bx->anglemax;
    if (amin < amax) {
        cd += 0.05 * fmax(0.0, amin - ang);
        cd += 0.05 * fmax(0.0, ang - amax);
    } else {
        cd += 0.05 * fmin(fmax(0.0, amin - ang), fmax(0.0, ang - amax));
    }
    return cd;
}

// Pass 1: bit b set <=> box bstart + b is still within 1.0 after the angle term (bcount <= 32, engine.BoxTable).
// The angle is only computed if some box survives the displacement / normal terms, and is returned for pass 2.
// (boxes are addressed as box_base + index * box_stride so that a padded shared-memory copy can be used)
__device__ __forceinline__ const fr3d_box_t *t_box_at(const char *box_base, size_t box_stride, int index) {
    return reinterpret_cast<const fr3d_box_t *>(box_base + (size_t)index * box_stride);
}

__device__ __forceinline__ unsigned t_boxes_live(const char *box_base, size_t box_stride, int bstart, int bcount, double dX,
                                                 double dY, double dZ, double nZ, double M11, double M01, double &ang) {
    unsigned m = 0;
#pragma unroll 2
    for (int b = 0; b < bcount; ++b)
        if (t_box_cd(t_box_at(box_base, box_stride, bstart + b), dX, dY, dZ, nZ) < 1.0) m |= 1u << b;   // NA:2580
    if (!m) return 0;
    ang = atan2_once(M11, M01) * 57.29577951308232 - 90.0;                           // NA:2594
    if (ang <= -90.0) ang += 360.0;
    unsigned live = 0;
    while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const fr3d_box_t *bx = t_box_at(box_base, box_stride, bstart + b);
        if (t_box_angle(bx, t_box_cd(bx, dX, dY, dZ, nZ), ang) < 1.0) live |= 1u << b;   // NA:2611
    }
    return live;
}

// Pass 2 over the surviving boxes: gap term, true / near classification and the choice (NA:2642-2759)
__device__ __forceinline__ void t_boxes_final(const char *box_base, size_t box_stride, int bstart, unsigned live, double dX,
                                              double dY, double dZ, double nZ, double ang, double gmax, double mind,
                                              int pb, TBp &out) {
    int best_match = -1, best_near = -1;
    int match_key = 0x7FFFFFFF;
    double near_cd = 1e300, match_cd = 0.0;
    int match_flags = 0, near_flags = 0;
    while (live) {
        const int b = __ffs(live) - 1;
        live &= live - 1;
        const fr3d_box_t *bx = t_box_at(box_base, box_stride, bstart + b);
        double cd = t_box_angle(bx, t_box_cd(bx, dX, dY, dZ, nZ), ang);
        const int flags = bx->flags;
        const double gapmax = bx->gapmax;
        if (gapmax > 0.1) cd += 4.0 * fmax(0.0, gmax - gapmax);      // NA:2643-2645
        bool one_hb = false;                                         // NA:2648-2652
        if ((flags & 2) && (pb == 1 || pb == 3)) one_hb = true;
        else if ((flags & 4) && pb == 3) one_hb = true;
        const bool starts_n = flags & 1;
        bool is_match = false, is_near = false;
        if (cd > 0) {
            if (cd < 1.0 && !starts_n && (mind < 4.1 || one_hb)) is_near = true;
        } else if (starts_n) {
            is_near = true;
        } else if (mind > 3.75 && !one_hb) {
            is_near = true;
        } else {
            is_match = true;
        }
        if (is_match) {                       // stable sort by (subcategory, len(name)): first smallest key wins
            const int key = bx->subcat * 64 + bx->name_len;
            if (key < match_key) { match_key = key; best_match = bstart + b; match_cd = cd; match_flags = flags; }
        } else if (is_near) {                 // nearest near match, earlier box on ties
            if (cd < near_cd) { near_cd = cd; best_near = bstart + b; near_flags = flags; }
        }
    }
    out.box = -1; out.near = 0; out.cd = 0.0;
    if (best_match >= 0) {
        out.box = best_match; out.cd = match_cd;
        out.near = ((match_cd > 0 && !(match_flags & 1)) ? 1 : 0) | ((match_flags & 1) ? 2 : 0);
    } else if (best_near >= 0) {
        out.box = best_near; out.cd = near_cd;
        out.near = ((near_cd > 0 && !(near_flags & 1)) ? 1 : 0) | ((near_flags & 1) ? 2 : 0);
    }
}

#ifndef FR3D_TPP_MIN_BLOCKS
#define FR3D_TPP_MIN_BLOCKS 3
#endif

// CATS_CT / BP_CT: the rule families compiled into the instantiation (the run-time mask is ANDed with it), so the
// kernel of one work list carries neither the code nor the registers of the other families.
template <int MIN_BLOCKS, uint32_t CATS_CT, bool BP_CT>
__global__ void __launch_bounds__(128, MIN_BLOCKS)
classify_pairs_tpp_kernel(fr3d_nts_t nts, const fr3d_box_t *__restrict__ boxes, const int32_t *__restrict__ box_index,
                          const int32_t *__restrict__ pair_i, const int32_t *__restrict__ pair_j,
                          const int32_t *__restrict__ pair_rank, int64_t pair_cap, uint32_t cats, int32_t index_offset,
                          fr3d_hit_t *__restrict__ hits, int64_t hit_cap, int64_t *counters,
                          const int32_t *__restrict__ worklist, int list_counter, int do_bp_rt,
                          const uint32_t *__restrict__ list_combos = nullptr) {
    cats &= CATS_CT;
    const bool do_bp = BP_CT && do_bp_rt;
    __shared__ double s_hull[FR3D_N_PARENTS * 7 * 3];        // a per-lane parent index would serialise constant loads
    for (int item = threadIdx.x; item < FR3D_N_PARENTS * 7 * 3; item += blockDim.x)
        s_hull[item] = (&c_tab.hull[0][0][0])[item];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // with a work list (classify_screen.cuh) the kernel walks its entries instead of the whole pair list
    long long n_pairs = worklist ? counters[list_counter] : counters[0];
    if (n_pairs > pair_cap) n_pairs = pair_cap;
    const long long stride = (long long)gridDim.x * blockDim.x;
    // whole warps iterate together so that the staging and the emission below can use warp collectives
    const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long rounds = (n_pairs + stride - 1) / stride;

#pragma unroll 1
    for (long long r = 0; r < rounds; ++r) {
        const long long p = first + r * stride;
        const bool active = p < n_pairs;
        if (!__any_sync(FULL, active)) continue;
        unsigned hmask = 0;
        int c_so12 = 0, c_so21 = 0, c_stack = 0, c_sr12 = 0, c_sr21 = 0, c_bph = 0, c_br = 0, c_bp = 0, bp_box = 0;
        double bp_cd = 0.0, bp_gap = 0.0;
        int n1 = 0, n2 = 0, rank = 0;
        if (active) {
            const long long q = worklist ? (long long)worklist[p] : p;
            n1 = pair_i[q]; n2 = pair_j[q]; rank = pair_rank[q];
        }
        uint32_t bm1 = 0, bm2 = 0, bb1 = 0, bb2 = 0;
        if (active) {
            bm1 = nts.base_mask[n1]; bm2 = nts.base_mask[n2];
            bb1 = nts.bb_mask[n1]; bb2 = nts.bb_mask[n2];
        }
        if (active) {
            const int par1 = (int)((bm1 >> 16) & 15u) - 1, par2 = (int)((bm2 >> 16) & 15u) - 1;
            const TFrame f1 = load_tframe(nts, n1), f2 = load_tframe(nts, n2);
            const double *base1 = nts.base_xyz + (size_t)n1 * FR3D_BASE_SLOTS * 3;
            const double *base2 = nts.base_xyz + (size_t)n2 * FR3D_BASE_SLOTS * 3;
            const double *bbp1 = nts.bb_xyz + (size_t)n1 * FR3D_BB_SLOTS * 3;
            const double *bbp2 = nts.bb_xyz + (size_t)n2 * FR3D_BB_SLOTS * 3;
            const double nZ = f1.u2 * f2.u2 + f1.u5 * f2.u5 + f1.u8 * f2.u8;      // M[2][2] in both orientations

            // ---- base-oxygen stacking, both directions (NA:758-778)
            if (cats & FR3D_CAT_SO) {
                const int a = t_so(f1, par1, bbp2, bb2);
                const int b = t_so(f2, par2, bbp1, bb1);
                if (a >= 0) { hmask |= 1u << FR3D_SLOT_SO12; c_so12 = a; }
                if (b >= 0) { hmask |= 1u << FR3D_SLOT_SO21; c_so21 = b; }
            }
            // ---- base-base stacking, part 1 (NA:1602-1702)
            int stack_code = -1;
            bool stack_both = false;
            if ((cats & FR3D_CAT_STACKING) && par1 >= 0 && par2 >= 0) {
                double z_2on1, z_1on2;
                const bool nt2on1 = t_overlap(f1, s_hull + par1 * 21, c_tab.hull_r2in[par1], base2, bm2, z_2on1);
                const bool nt1on2 = t_overlap(f2, s_hull + par2 * 21, c_tab.hull_r2in[par2], base1, bm1, z_1on2);
                if ((nt2on1 || nt1on2) && !(fabs(nZ) < 0.5)) {
                    const bool reverse = nt1on2 && !nt2on1;
                    const double mvd = reverse ? z_1on2 : z_2on1;
                    int code = -1;
                    if (mvd > 0) { if (nZ > 0) code = 0; else if (nZ < 0) code = 2; }
                    else if (mvd < 0) { if (nZ > 0) code = 1; else if (nZ < 0) code = 3; }
                    if (reverse && code >= 0 && code < 2) code ^= 1;
                    stack_code = code;
                    stack_both = nt2on1 && nt1on2;
                }
            }
            // ---- sugar-ribose, both directions (NA:791-804)
            if ((cats & FR3D_CAT_SUGAR_RIBOSE) && par1 >= 0 && par2 >= 0 && par1 != 4 && par2 != 4 &&
                (bb1 >> 6 & 1u) && (bb2 >> 6 & 1u)) {
                const double *A = bbp1, *B = bbp2;
                const double o0 = A[18] - B[18], o1 = A[19] - B[19], o2 = A[20] - B[20];
                const double oo = o0 * o0 + o1 * o1 + o2 * o2;
                if (oo < 14.4401 && !(sqrt(oo) > 3.8)) {
                    const int r12 = t_sr_tail(bbp1, bbp2, base1, bm1, bb1, bb2, par1, f1.u2, f1.u5, f1.u8);
                    const int r21 = t_sr_tail(bbp2, bbp1, base2, bm2, bb2, bb1, par2, f2.u2, f2.u5, f2.u8);
                    if (r12 >= 0) { hmask |= 1u << FR3D_SLOT_SR12; c_sr12 = r12; }
                    if (r21 >= 0) { hmask |= 1u << FR3D_SLOT_SR21; c_sr21 = r21; }
                }
            }
            // ---- base-phosphate / base-ribose, nt1 base -> nt2 backbone (NA:807-829)
            // (list_combos: the candidate (site, oxygen) combinations of each listed pair, from the screen)
            const unsigned combos = list_combos ? list_combos[p] : 0xFFFFFFFFu;
            if ((cats & FR3D_CAT_BACKBONE) && par1 >= 0 && combos &&
                (list_combos || t_backbone_possible(par1, base1, bbp2, bm1, bb2))) {
                const int sel = t_backbone(f1, par1, base1, bbp2, bm1, bb2, combos);
                if ((sel & 0xFF) != 0xFF) { hmask |= 1u << FR3D_SLOT_BPH; c_bph = sel & 0xFF; }
                if (((sel >> 8) & 0xFF) != 0xFF) { hmask |= 1u << FR3D_SLOT_BR; c_br = (sel >> 8) & 0xFF; }
            }
            // ---- coplanar + basepair (NA:834-1013)
            const bool gly_ok = do_bp && (bm1 & 1u) && (bm2 & 1u) && par1 >= 0 && par2 >= 0;
            unsigned ori_ok = 0, bp_live = 0;
            bool cp_frames = false;
            double gd0 = 0, gd1 = 0, gd2 = 0;
            unsigned bx_live[2] = {0u, 0u};                  // surviving boxes, angle and displacement per orientation
            int bx_start[2] = {0, 0};
            double bx_ang[2] = {0.0, 0.0}, bx_d[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
            if (gly_ok) {
                if (c_tab.combo_ok[par1 * FR3D_N_PARENTS + par2]) ori_ok |= 1u;
                if (c_tab.combo_ok[par2 * FR3D_N_PARENTS + par1]) ori_ok |= 2u;
                if ((cats & FR3D_CAT_COPLANAR) && ori_ok && fabs(nZ) > 0.7757) {
                    const double e0 = f1.c0 - f2.c0, e1 = f1.c1 - f2.c1, e2 = f1.c2 - f2.c2;
                    const double lim = (0.3381 * 0.3381) * (e0 * e0 + e1 * e1 + e2 * e2);
                    const double d1 = e0 * f1.u2 + e1 * f1.u5 + e2 * f1.u8;
                    const double d2 = e0 * f2.u2 + e1 * f2.u5 + e2 * f2.u8;
                    cp_frames = d1 * d1 < lim && d2 * d2 < lim;
                }
                const double *g1 = base1, *g2 = base2;
                gd0 = g2[0] - g1[0]; gd1 = g2[1] - g1[1]; gd2 = g2[2] - g1[2];
#pragma unroll
                for (int ori = 0; ori < 2; ++ori) {
                    if (!((ori_ok >> ori) & 1u)) continue;
                    const TFrame &fa = ori ? f2 : f1;
                    const TFrame &fb = ori ? f1 : f2;
                    const double s = ori ? -1.0 : 1.0;
                    const double v0 = s * gd0, v1 = s * gd1, v2 = s * gd2;
                    const double dZ = v0 * fa.u2 + v1 * fa.u5 + v2 * fa.u8;
                    if (fabs(dZ) > 3.6) continue;
                    const double dX = v0 * fa.u0 + v1 * fa.u3 + v2 * fa.u6;
                    const double dY = v0 * fa.u1 + v1 * fa.u4 + v2 * fa.u7;
                    const int combo = ori ? par2 * FR3D_N_PARENTS + par1 : par1 * FR3D_N_PARENTS + par2;
                    const int sg = nZ > 0 ? 0 : 1;
                    const int bstart = box_index[(combo * 2 + sg) * 2 + 0];
                    const int bcount = box_index[(combo * 2 + sg) * 2 + 1];
                    const double M11 = fa.u1 * fb.u1 + fa.u4 * fb.u4 + fa.u7 * fb.u7;
                    const double M01 = fa.u0 * fb.u1 + fa.u3 * fb.u4 + fa.u6 * fb.u7;
                    bx_live[ori] = t_boxes_live(reinterpret_cast<const char *>(boxes), sizeof(fr3d_box_t), bstart, bcount, dX, dY, dZ, nZ,
                                                M11, M01, bx_ang[ori]);
                    bx_start[ori] = bstart;
                    bx_d[ori][0] = dX; bx_d[ori][1] = dY; bx_d[ori][2] = dZ;
                    if (bx_live[ori]) bp_live |= 1u << ori;
                }
            }
            // ---- gaps / minimum distance, only when something can still use them
            double gap12 = 100.0, gap21 = 100.0, mind = 9999.0;
            if (stack_code >= 0 || bp_live || cp_frames) {
                uint32_t tw1 = 0, tw2 = 0;
                if ((bm1 >> 22) & 1u) tw1 = (uint32_t)nts.meta[(size_t)n1 * FR3D_META_INTS + 7] >> 16;
                if ((bm2 >> 22) & 1u) tw2 = (uint32_t)nts.meta[(size_t)n2 * FR3D_META_INTS + 7] >> 16;
                if (bp_live || cp_frames) {
                    gap12 = t_gap(f1, f2, par2, base2, bm2, tw2);
                    gap21 = t_gap(f2, f1, par1, base1, bm1, tw1);
                }
                const bool cp_gap = cp_frames && (((ori_ok & 1u) && gap12 < 1.5179) || ((ori_ok & 2u) && gap21 < 1.5179));
                if (stack_code >= 0 || bp_live || cp_gap)
                    mind = t_mind(f1, f2, par1, par2, base1, base2, bm1, bm2, tw1, tw2);
                if (stack_code >= 0 && !(fabs(mind) < 1.0)) {                 // NA:1707-1718
                    int near = 1;
                    if (fabs(mind) < 4.0 && fabs(mind) > 1.0 && fabs(nZ) > 0.6 && stack_both) near = 0;
                    hmask |= 1u << FR3D_SLOT_STACK;
                    c_stack = stack_code | (near << 2);
                }
                if (cp_gap) {                                                 // NA:2080-2096
                    const bool both_std = (bm1 >> 20 & 1u) && (bm2 >> 20 & 1u);
                    if (mind < 3.4589 && !(mind >= 2.4589 && both_std)) hmask |= 1u << FR3D_SLOT_CP;
                }
                if (bp_live) {
                    const double gmax = fmax(gap12, gap21);
                    TBp res[2];
                    res[0].box = -1; res[1].box = -1; res[0].near = res[1].near = 0; res[0].cd = res[1].cd = 0.0;
#pragma unroll
                    for (int ori = 0; ori < 2; ++ori) {
                        if (!((bp_live >> ori) & 1u)) continue;
                        t_boxes_final(reinterpret_cast<const char *>(boxes), sizeof(fr3d_box_t), bx_start[ori], bx_live[ori], bx_d[ori][0], bx_d[ori][1], bx_d[ori][2], nZ,
                                      bx_ang[ori], gmax, mind, ori ? par1 : par2, res[ori]);
                    }
                    int use = -1;                                             // NA:935-957
                    if (res[0].box >= 0 && res[1].box >= 0) {
                        const bool n0 = res[0].near != 0, n1_ = res[1].near != 0;
                        const bool same = ((res[0].near & 1) == (res[1].near & 1)) &&
                                          (boxes[res[0].box].rev_lower_id == boxes[res[1].box].lower_id);
                        if (same) use = 0;
                        else if (n0 && n1_) use = (res[0].cd < res[1].cd) ? 0 : 1;
                        else if (n1_) use = 0;
                        else if (n0) use = 1;
                        else use = 0;
                    } else if (res[0].box >= 0) use = 0;
                    else if (res[1].box >= 0) use = 1;
                    if (use >= 0) {
                        hmask |= 1u << FR3D_SLOT_BP;
                        const TBp &u = use ? res[1] : res[0];
                        c_bp = (u.near & 1) | (use << 1);
                        bp_box = u.box;
                        bp_cd = u.cd;
                        bp_gap = gmax;
                    }
                }
            }
        }
        // ---- emit: warp-aggregated allocation (one atomicAdd per warp), then each thread writes its hits
        const int cnt = __popc(hmask);
        int incl = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, off);
            if (lane >= off) incl += t;
        }
        const int total = __shfl_sync(FULL, incl, 31);
        if (total) {
            long long base = 0;
            if (lane == 31) base = (long long)atomicAdd((unsigned long long *)&counters[1], (unsigned long long)total);
            base = __shfl_sync(FULL, base, 31);
            long long pos = base + (incl - cnt);
            unsigned m = hmask;
            while (m) {
                const int s = __ffs(m) - 1;
                m &= m - 1;
                if (pos < hit_cap) {
                    int code = 0;
                    switch (s) {
                        case FR3D_SLOT_SO12: code = c_so12; break;
                        case FR3D_SLOT_SO21: code = c_so21; break;
                        case FR3D_SLOT_STACK: code = c_stack; break;
                        case FR3D_SLOT_SR12: code = c_sr12; break;
                        case FR3D_SLOT_SR21: code = c_sr21; break;
                        case FR3D_SLOT_BPH: code = c_bph; break;
                        case FR3D_SLOT_BR: code = c_br; break;
                        case FR3D_SLOT_BP: code = c_bp; break;
                        default: break;
                    }
                    const bool isbp = s == FR3D_SLOT_BP;
                    const double cdv = isbp ? bp_cd : 0.0, gv = isbp ? bp_gap : 0.0;
                    const uint32_t w3 = (uint32_t)s | ((uint32_t)code << 8) | ((uint32_t)(isbp ? bp_box : 0) << 16);
                    uint4 *dst = reinterpret_cast<uint4 *>(hits + pos);
                    dst[0] = make_uint4((uint32_t)(n1 + index_offset), (uint32_t)(n2 + index_offset),
                                        (uint32_t)(rank + 16 * index_offset), w3);
                    dst[1] = make_uint4((uint32_t)__double2loint(cdv), (uint32_t)__double2hiint(cdv),
                                        (uint32_t)__double2loint(gv), (uint32_t)__double2hiint(gv));
                }
                ++pos;
            }
        }
    }
}

}  // namespace fr3d
