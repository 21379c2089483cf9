// K4a: thread-per-pair screen.
//
// The reference hands every pair with 2 A <= d <= 12 A to all of its rule functions; about two thirds of
// those pairs end with no annotation at all.  ncu showed (profiles/README.md) that a warp-per-pair kernel
// spends most of its ~950 instructions per pair on warp-uniform bookkeeping just to find that out, and
// that a pure thread-per-pair kernel is divergence bound (7.8 of 32 lanes active) because the rare, long
// tails serialise the warp.  So the work is split: this kernel evaluates, one THREAD per pair and fully
// converged, the cheap NECESSARY condition of every rule family (the same predicates the warp-per-pair
// kernel votes on) and compacts the pairs that can still produce something into two work lists:
//   detail list  -> classify_pairs_tpp_kernel   (sO, sugar-ribose, base-phosphate / base-ribose; thread per pair)
//   geometry list-> classify_pairs_pm_kernel    (stacking, coplanar, basepair; warp per pair, shared-memory staging)
// Every predicate is a superset of "this family yields a hit" (each rule is a conjunction), so results are
// unchanged; tests/test_gpu_parity.py compares complete annotation sets with the reference.
#pragma once

#include "classify_tpp.cuh"

namespace fr3d {

#define SCR_SO12 1u
#define SCR_SO21 2u
#define SCR_STACK 4u
#define SCR_SR 8u
#define SCR_BPH 16u
#define SCR_CP 32u
#define SCR_BP 64u
#define SCR_DETAIL (SCR_SO12 | SCR_SO21 | SCR_SR | SCR_BPH)
#define SCR_GEOM (SCR_STACK | SCR_CP | SCR_BP)

// any base atom of nt_b inside the hull prism of nt_a?  (check_convex_hull_atoms, NA:1475-1528; every hull
// lies inside the disc x^2 + y^2 < 16, tests/test_abi.py)
__device__ __forceinline__ bool s_any_inside(const fr3d_nts_t &nts, const TFrame &fa, int par_a, int nb, uint32_t bmb) {
    const double *q = nts.base_xyz + (size_t)nb * FR3D_BASE_SLOTS * 3;
    bool any = false;
#pragma unroll 4
    for (int a = 0; a < FR3D_BASE_SLOTS; ++a) {
        const double v0 = q[a * 3] - fa.c0, v1 = q[a * 3 + 1] - fa.c1, v2 = q[a * 3 + 2] - fa.c2;
        const double z = v0 * fa.u2 + v1 * fa.u5 + v2 * fa.u8;
        const double x = v0 * fa.u0 + v1 * fa.u3 + v2 * fa.u6;
        const double y = v0 * fa.u1 + v1 * fa.u4 + v2 * fa.u7;
        if (((bmb >> a) & 1u) && fabs(z) < 4.5 && (x * x + y * y) < 16.0) {
            bool in = true;
#pragma unroll
            for (int k = 0; k < 7; ++k) in = in && left_of(c_tab.hull[par_a][k], x, y);
            any = any || in;
        }
    }
    return any;
}

// any backbone oxygen of nt_b inside the slab / disc that contains every ring of nt_a's base?
__device__ __forceinline__ bool s_so_possible(const fr3d_nts_t &nts, const TFrame &fa, int pa, int nb, uint32_t bbb) {
    if (pa < 0 || pa >= FR3D_N_PARENTS) return false;
    const double *B = nts.bb_xyz + (size_t)nb * FR3D_BB_SLOTS * 3;
    bool any = false;
#pragma unroll
    for (int o = 0; o < 6; ++o) {
        const int oslot = (0x213456 >> (4 * o)) & 7;
        const double v0 = B[oslot * 3] - fa.c0, v1 = B[oslot * 3 + 1] - fa.c1, v2 = B[oslot * 3 + 2] - fa.c2;
        const double z = v0 * fa.u2 + v1 * fa.u5 + v2 * fa.u8;
        const double x = v0 * fa.u0 + v1 * fa.u3 + v2 * fa.u6;
        const double y = v0 * fa.u1 + v1 * fa.u4 + v2 * fa.u7;
        any = any || (((bbb >> oslot) & 1u) && fabs(z) < 3.6 && (x * x + y * y) < 25.0);
    }
    return any;
}

__global__ void __launch_bounds__(128)
classify_screen_kernel(fr3d_nts_t nts, const fr3d_box_t *__restrict__ boxes, const int32_t *__restrict__ box_index,
                       const int32_t *__restrict__ pair_i, const int32_t *__restrict__ pair_j, int64_t pair_cap,
                       uint32_t cats, int32_t *__restrict__ wl_detail, int32_t *__restrict__ wl_geom,
                       int64_t *counters) {
    const int lane = threadIdx.x & 31;
    long long n_pairs = counters[0];
    if (n_pairs > pair_cap) n_pairs = pair_cap;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long rounds = (n_pairs + stride - 1) / stride;
#pragma unroll 1
    for (long long r = 0; r < rounds; ++r) {
        const long long p = first + r * stride;
        unsigned fl = 0;
        if (p < n_pairs) {
            const int n1 = pair_i[p], n2 = pair_j[p];
            const uint32_t bm1 = nts.base_mask[n1], bm2 = nts.base_mask[n2];
            const uint32_t bb1 = nts.bb_mask[n1], bb2 = nts.bb_mask[n2];
            const int par1 = (int)((bm1 >> 16) & 15u) - 1, par2 = (int)((bm2 >> 16) & 15u) - 1;
            const TFrame f1 = load_tframe(nts, n1), f2 = load_tframe(nts, n2);
            if (cats & FR3D_CAT_SO) {
                if (s_so_possible(nts, f1, par1, n2, bb2)) fl |= SCR_SO12;
                if (s_so_possible(nts, f2, par2, n1, bb1)) fl |= SCR_SO21;
            }
            if ((cats & FR3D_CAT_STACKING) && par1 >= 0 && par2 >= 0) {
                if (s_any_inside(nts, f1, par1, n2, bm2) || s_any_inside(nts, f2, par2, n1, bm1)) fl |= SCR_STACK;
            }
            if ((cats & FR3D_CAT_SUGAR_RIBOSE) && par1 >= 0 && par2 >= 0 && par1 != 4 && par2 != 4 &&
                (bb1 >> 6 & 1u) && (bb2 >> 6 & 1u)) {
                const double *A = nts.bb_xyz + (size_t)n1 * FR3D_BB_SLOTS * 3;
                const double *B = nts.bb_xyz + (size_t)n2 * FR3D_BB_SLOTS * 3;
                const double o0 = A[18] - B[18], o1 = A[19] - B[19], o2 = A[20] - B[20];
                if (o0 * o0 + o1 * o1 + o2 * o2 < 14.4401) fl |= SCR_SR;
            }
            if ((cats & FR3D_CAT_BACKBONE) && par1 >= 0 && t_backbone_possible(nts, par1, n1, n2, bm1, bb2)) fl |= SCR_BPH;
            if ((bm1 & 1u) && (bm2 & 1u) && par1 >= 0 && par2 >= 0) {
                unsigned ori_ok = 0;
                if (c_tab.combo_ok[par1 * FR3D_N_PARENTS + par2]) ori_ok |= 1u;
                if (c_tab.combo_ok[par2 * FR3D_N_PARENTS + par1]) ori_ok |= 2u;
                const double nZ = f1.u2 * f2.u2 + f1.u5 * f2.u5 + f1.u8 * f2.u8;
                if ((cats & FR3D_CAT_COPLANAR) && ori_ok && fabs(nZ) > 0.7757) {
                    const double e0 = f1.c0 - f2.c0, e1 = f1.c1 - f2.c1, e2 = f1.c2 - f2.c2;
                    const double lim = (0.3381 * 0.3381) * (e0 * e0 + e1 * e1 + e2 * e2);
                    const double d1 = e0 * f1.u2 + e1 * f1.u5 + e2 * f1.u8;
                    const double d2 = e0 * f2.u2 + e1 * f2.u5 + e2 * f2.u8;
                    if (d1 * d1 < lim && d2 * d2 < lim) fl |= SCR_CP;
                }
                const double *g1 = nts.base_xyz + (size_t)n1 * FR3D_BASE_SLOTS * 3;
                const double *g2 = nts.base_xyz + (size_t)n2 * FR3D_BASE_SLOTS * 3;
                const double gd0 = g2[0] - g1[0], gd1 = g2[1] - g1[1], gd2 = g2[2] - g1[2];
#pragma unroll 1
                for (int ori = 0; ori < 2; ++ori) {
                    if (!((ori_ok >> ori) & 1u) || (fl & SCR_BP)) continue;
                    const TFrame &fa = ori ? f2 : f1;
                    const TFrame &fb = ori ? f1 : f2;
                    const double s = ori ? -1.0 : 1.0;
                    const double v0 = s * gd0, v1 = s * gd1, v2 = s * gd2;
                    const double dZ = v0 * fa.u2 + v1 * fa.u5 + v2 * fa.u8;
                    if (fabs(dZ) > 3.6) continue;                                  // NA:2380
                    const double dX = v0 * fa.u0 + v1 * fa.u3 + v2 * fa.u6;
                    const double dY = v0 * fa.u1 + v1 * fa.u4 + v2 * fa.u7;
                    const int combo = ori ? par2 * FR3D_N_PARENTS + par1 : par1 * FR3D_N_PARENTS + par2;
                    const int sg = nZ > 0 ? 0 : 1;
                    const int bstart = box_index[(combo * 2 + sg) * 2 + 0];
                    const int bcount = box_index[(combo * 2 + sg) * 2 + 1];
                    // some box within 1.0 before the angle term (NA:2546-2581); the angle and gap terms only add
                    bool any = false;
#pragma unroll 1
                    for (int b = 0; b < bcount && !any; ++b) {
                        const fr3d_box_t *bx = boxes + bstart + b;
                        double cd = fmax(0.0, bx->xmin - dX) + fmax(0.0, dX - bx->xmax) + fmax(0.0, bx->ymin - dY) +
                                    fmax(0.0, dY - bx->ymax) + fmax(0.0, bx->zmin - dZ) + fmax(0.0, dZ - bx->zmax) +
                                    3.0 * (fmax(0.0, bx->normalmin - nZ) + fmax(0.0, nZ - bx->normalmax));
                        any = cd < 1.0 + 1e-9;         // summation order differs from the reference's: keep a margin
                    }
                    if (any) fl |= SCR_BP;
                    (void)fb;
                }
            }
        }
        // warp-aggregated appends to the two work lists
        const unsigned bd = __ballot_sync(FULL, (fl & SCR_DETAIL) != 0);
        const unsigned bg = __ballot_sync(FULL, (fl & SCR_GEOM) != 0);
        if (bd) {
            long long base = 0;
            if (lane == 0) base = (long long)atomicAdd((unsigned long long *)&counters[4], (unsigned long long)__popc(bd));
            base = __shfl_sync(FULL, base, 0);
            if (fl & SCR_DETAIL) wl_detail[base + __popc(bd & ((1u << lane) - 1u))] = (int32_t)p;
        }
        if (bg) {
            long long base = 0;
            if (lane == 0) base = (long long)atomicAdd((unsigned long long *)&counters[5], (unsigned long long)__popc(bg));
            base = __shfl_sync(FULL, base, 0);
            if (fl & SCR_GEOM) wl_geom[base + __popc(bg & ((1u << lane) - 1u))] = (int32_t)p;
        }
    }
}

}  // namespace fr3d
