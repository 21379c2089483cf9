// K4a: thread-per-pair screen.
//
// The reference hands every pair with 2 A <= d <= 12 A to all of its rule functions; about two thirds of
// those pairs end with no annotation at all.  ncu showed (profiles/README.md) that a warp-per-pair kernel
// spends most of its ~950 instructions per pair on warp-uniform bookkeeping just to find that out, and
// that a pure thread-per-pair kernel is divergence bound (7.8 of 32 lanes active) because the rare, long
// tails serialise the warp.  So the work is split: this kernel evaluates, one THREAD per pair and fully
// converged, the cheap NECESSARY condition of every rule family (the same predicates the warp-per-pair
// kernel votes on) and compacts the pairs that can still produce something into four work lists:
//   light list    -> classify_pairs_tpp_kernel  (sO, sugar-ribose)
//   phosphate list-> classify_pairs_tpp_kernel  (base-phosphate / base-ribose)
//   stacking list -> classify_list2_kernel      (base stacking)
//   pairing list  -> classify_list2_kernel      (coplanar, basepair)
// Each list is homogeneous - nearly every pair on it runs the same rule to the end - so a thread-per-pair kernel
// stays converged on it.
// Every predicate is a superset of "this family yields a hit" (each rule is a conjunction), so results are
// unchanged; tests/test_gpu_parity.py compares complete annotation sets with the reference.
#pragma once

#include "classify_pm.cuh"
#include "classify_tpp.cuh"

namespace fr3d {

#define SCR_SO12 1u
#define SCR_SO21 2u
#define SCR_STACK 4u
#define SCR_SR 8u
#define SCR_BPH 16u
#define SCR_CP 32u
#define SCR_BP 64u

// Shared-memory row of one nt2 record (doubles), staged in two phases that re-use the same buffer so that twice as
// many warps fit on an SM:
//   phase A: centre 0..2, rotation 3..11, backbone slots 1..6 (OP1 OP2 O5' O4' O3' O2') 12..29, x and y of slot 9
//            (previous O3') 30..31 - exactly one copy instruction per row; z of slot 9 is loaded by the owning
//            lane directly                     -> sO, sugar-ribose, base-phosphate screens
//   phase B: the 16 base points 0..47          -> stacking, coplanar / basepair screens
// The stride is odd so that "lane k reads its own row" is bank-conflict free for 64-bit accesses.
constexpr int SCR_ROW_A = 32;
constexpr int SCR_ROW_B = 48;
constexpr int SCR_STRIDE = 49;
#define SCR_BB_OFF(slot) ((slot) == 9 ? 30 : 9 + (slot) * 3)
#ifndef FR3D_SCR_WARPS
#define FR3D_SCR_WARPS 16
#endif
constexpr int SCR_WARPS = FR3D_SCR_WARPS;
constexpr int SCR_UBOX = FR3D_N_PARENTS * FR3D_N_PARENTS * 2;      // (combo, sign of nZ) entries
constexpr int SCR_MAX_BOXES = 256;                                 // cutoff boxes kept in shared memory (else global)
constexpr int SCR_BOX_STRIDE = 11;                                 // 10 floats per box; odd: lane b reads box b conflict free
constexpr int SCR_HULL = FR3D_N_PARENTS * 7 * 3;
constexpr int SCR_FIXED = SCR_UBOX * 8 + (SCR_MAX_BOXES * SCR_BOX_STRIDE + 1) / 2 + SCR_HULL + SCR_UBOX;   // doubles
constexpr size_t SCR_SMEM = sizeof(double) * (SCR_FIXED + SCR_WARPS * 32 * SCR_STRIDE);

// which atoms (bit per slot) of the 16 points q[0..47] lie in the cylinder |z| < 4.5, x^2 + y^2 < r2max of frame
// fa (r2max = c_tab.hull_r2max[parent of fa]: the hull lies inside that disc, tests/test_abi.py); converged,
// branch free
__device__ __forceinline__ unsigned s_cylinder_mask(const TFrame &fa, const double *q, double r2max) {
    unsigned cand = 0;
#pragma unroll
    for (int a = 0; a < FR3D_BASE_SLOTS; ++a) {
        const double v0 = q[a * 3] - fa.c0, v1 = q[a * 3 + 1] - fa.c1, v2 = q[a * 3 + 2] - fa.c2;
        const double z = v0 * fa.u2 + v1 * fa.u5 + v2 * fa.u8;
        const double r2 = (v0 * v0 + v1 * v1 + v2 * v2) - z * z;        // x^2 + y^2 (U is orthonormal)
        if (fabs(z) < 4.5 && r2 < r2max) cand |= 1u << a;
    }
    return cand;
}

// any of the candidate atoms inside the hull prism of fa's base?  (check_convex_hull_atoms, NA:1475-1528)
__device__ __forceinline__ bool s_hull_any(const TFrame &fa, const double *hull, const double *q, unsigned cand) {
    bool any = false;
    while (cand && !any) {
        const int a = __ffs(cand) - 1;
        cand &= cand - 1;
        double x, y, z;
        tproject(fa, q[a * 3], q[a * 3 + 1], q[a * 3 + 2], x, y, z);
        bool in = fabs(z) < 4.5;
#pragma unroll
        for (int k = 0; k < 7; ++k) in = in && left_of(hull + k * 3, x, y);
        any = in;
    }
    return any;
}

// any backbone oxygen (slots 1..6 of the row B) inside the slab / disc (r2max = c_tab.so_r2max[pa]) that contains
// every ring and near-ring ellipse of fa's base?
__device__ __forceinline__ bool s_so_possible(const TFrame &fa, int pa, const double *B, uint32_t bbb, double r2max) {
    if (pa < 0 || pa >= FR3D_N_PARENTS) return false;
    bool any = false;
#pragma unroll
    for (int oslot = 1; oslot <= 6; ++oslot) {
        const double v0 = B[oslot * 3] - fa.c0, v1 = B[oslot * 3 + 1] - fa.c1, v2 = B[oslot * 3 + 2] - fa.c2;
        const double z = v0 * fa.u2 + v1 * fa.u5 + v2 * fa.u8;
        const double r2 = (v0 * v0 + v1 * v1 + v2 * v2) - z * z;
        any = any || (((bbb >> oslot) & 1u) && fabs(z) < 3.6 && r2 < r2max);
    }
    return any;
}

__global__ void __launch_bounds__(SCR_WARPS * 32)
classify_screen_kernel(fr3d_nts_t nts, const fr3d_box_t *__restrict__ boxes, const int32_t *__restrict__ box_index,
                       const int32_t *__restrict__ pair_i, const int32_t *__restrict__ pair_j, int64_t pair_cap,
                       uint32_t cats, int32_t *__restrict__ wl_light, int32_t *__restrict__ wl_stack,
                       int32_t *__restrict__ wl_pair, int32_t *__restrict__ wl_bph, int64_t *counters) {
    extern __shared__ double scr_smem[];
    double *ubox = scr_smem;                                                  // [SCR_UBOX][8] union boxes
    float *sbox = reinterpret_cast<float *>(ubox + SCR_UBOX * 8);             // [SCR_MAX_BOXES][SCR_BOX_STRIDE] mid, half
    double *shull = ubox + SCR_UBOX * 8 + (SCR_MAX_BOXES * SCR_BOX_STRIDE + 1) / 2;   // [parents][7][3]
    int *sidx = reinterpret_cast<int *>(shull + SCR_HULL);                    // [SCR_UBOX][2] start, count
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *rows = scr_smem + SCR_FIXED + warp * 32 * SCR_STRIDE;
    __shared__ int s_nboxes;
    __shared__ double s_r2[2 * FR3D_N_PARENTS + 2];       // [parent + 1]: sO disc, then hull disc (index 0: no parent)
    if (threadIdx.x == 0) s_nboxes = 0;
    if (threadIdx.x <= FR3D_N_PARENTS) {
        const int p = (int)threadIdx.x - 1;
        s_r2[threadIdx.x] = p < 0 ? 0.0 : c_tab.so_r2max[p] + 1e-9;
        s_r2[FR3D_N_PARENTS + 1 + threadIdx.x] = p < 0 ? 0.0 : c_tab.hull_r2max[p] + 1e-9;
    }
    __syncthreads();
    // prologue: the union of the cutoff boxes of every (combination, sign) entry.  A pair whose L1 excess over the
    // union box already reaches 1 cannot come within 1 of any single box (each box's excess is at least the union's).
    for (int item = threadIdx.x; item < SCR_UBOX * 8; item += blockDim.x) {
        const int e = item >> 3, q = item & 7;
        const int bstart = box_index[e * 2], bcount = box_index[e * 2 + 1];
        const bool is_min = !(q & 1);
        double acc = is_min ? 1e300 : -1e300;
        for (int b = 0; b < bcount; ++b) {
            const double v = reinterpret_cast<const double *>(boxes + bstart + b)[q];
            acc = is_min ? fmin(acc, v) : fmax(acc, v);
        }
        ubox[item] = acc;
        if (q == 0) { sidx[e * 2] = bstart; sidx[e * 2 + 1] = bcount; if (bcount > 0) atomicMax(&s_nboxes, bstart + bcount); }
    }
    for (int item = threadIdx.x; item < SCR_HULL; item += blockDim.x) shull[item] = (&c_tab.hull[0][0][0])[item];
    __syncthreads();
    // Cutoff boxes as single-precision (mid, half-width) pairs: the L1 excess of v over [lo, hi] is
    // max(0, |v - mid| - half).  Single precision is enough for a screen: the acceptance margin below (1e-3) is far
    // above its rounding error for displacements of a few Angstrom, so the screen stays a superset.
    const bool boxes_in_smem = s_nboxes <= SCR_MAX_BOXES;
    int max_count = 0;
    for (int e = 0; e < SCR_UBOX; ++e) max_count = max(max_count, sidx[e * 2 + 1]);
    const bool two_tasks = boxes_in_smem && max_count <= 16;        // one (pair, orientation) per half warp
    if (boxes_in_smem)
        for (int item = threadIdx.x; item < s_nboxes * 4; item += blockDim.x) {
            const double *bx = reinterpret_cast<const double *>(boxes + (item >> 2));
            const double lo = bx[(item & 3) * 2], hi = bx[(item & 3) * 2 + 1];
            sbox[(item >> 2) * SCR_BOX_STRIDE + (item & 3) * 2] = (float)(0.5 * (lo + hi));
            sbox[(item >> 2) * SCR_BOX_STRIDE + (item & 3) * 2 + 1] = (float)(0.5 * (hi - lo));
            if ((item & 3) == 0) {                                  // the in-plane angle range (may wrap around)
                sbox[(item >> 2) * SCR_BOX_STRIDE + 8] = (float)bx[8];
                sbox[(item >> 2) * SCR_BOX_STRIDE + 9] = (float)bx[9];
            }
        }
    __syncthreads();
    // per-lane sources of the copies that move one row (phase A: 32 doubles, phase B: 32 + 16)
    const double *src_a0, *src_b0 = nts.base_xyz + lane, *src_b1 = nts.base_xyz + 32 + lane;
    int mult_a0;
    if (lane < 3) { src_a0 = nts.center + lane; mult_a0 = 3; }
    else if (lane < 12) { src_a0 = nts.rot + (lane - 3); mult_a0 = 9; }
    else if (lane < 30) { src_a0 = nts.bb_xyz + (lane - 12 + 3); mult_a0 = FR3D_BB_SLOTS * 3; }
    else { src_a0 = nts.bb_xyz + (lane - 30 + 27); mult_a0 = FR3D_BB_SLOTS * 3; }
    long long n_pairs = counters[0];
    if (n_pairs > pair_cap) n_pairs = pair_cap;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long rounds = (n_pairs + stride - 1) / stride;
    // the nt indices of a round are fetched one round ahead, so the staging below never waits for them
    int nx1 = first < n_pairs ? pair_i[first] : 0, nx2 = first < n_pairs ? pair_j[first] : 0;
#pragma unroll 1
    for (long long r = 0; r < rounds; ++r) {
        const long long p = first + r * stride;
        const bool valid = p < n_pairs;
        const int n1 = nx1, n2 = nx2;
        nx1 = p + stride < n_pairs ? pair_i[p + stride] : 0;
        nx2 = p + stride < n_pairs ? pair_j[p + stride] : 0;
        if (!__any_sync(FULL, valid)) continue;
        // phase A: stage frame + backbone oxygens of the 32 nt2 records with coalesced asynchronous copies
        // (row k belongs to lane k)
        __syncwarp();
#pragma unroll 4
        for (int k = 0; k < 32; ++k) {
            const int nb = __shfl_sync(FULL, n2, k);
            double *dst = rows + k * SCR_STRIDE + lane;
            cp_async8(dst, src_a0 + (size_t)nb * mult_a0);
        }
        const double prev_o3_z = nts.bb_xyz[(size_t)n2 * (FR3D_BB_SLOTS * 3) + 29];
        unsigned fl = 0, live = 0;
        double dX0 = 0, dY0 = 0, dZ0 = 0, dX1 = 0, dY1 = 0, dZ1 = 0, nZ = 0;
        float ang0 = 0.f, ang1 = 0.f;
        int entry0 = 0, entry1 = 0;
        const uint32_t bm1 = nts.base_mask[n1], bm2 = nts.base_mask[n2];
        const uint32_t bb1 = nts.bb_mask[n1], bb2 = nts.bb_mask[n2];
        const int par1 = (int)((bm1 >> 16) & 15u) - 1, par2 = (int)((bm2 >> 16) & 15u) - 1;
        const TFrame f1 = load_tframe(nts, n1);
        const double *A_base = nts.base_xyz + (size_t)n1 * FR3D_BASE_SLOTS * 3;     // nt1: mostly warp-uniform addresses
        const double *A_bb = nts.bb_xyz + (size_t)n1 * FR3D_BB_SLOTS * 3;
        // pull nt1's lines towards L1 now, while the staging copies are in flight: the first touches further down
        // (site atoms in a serial loop, the cylinder pass) otherwise each wait for L2
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A_base));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A_base + 16));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A_base + 32));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A_base + 47));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A_bb));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A_bb + 16));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A_bb + 29));
        {   // ... and the next round's nt1 frame and masks (their indices are already here)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(nts.center + (size_t)nx1 * 3));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(nts.rot + (size_t)nx1 * 9));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(nts.rot + (size_t)nx1 * 9 + 8));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(nts.base_mask + nx1));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(nts.base_mask + nx2));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(nts.bb_mask + nx1));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(nts.bb_mask + nx2));
        }
        cp_async_wait_all();
        __syncwarp();
        const double *R = rows + lane * SCR_STRIDE;
        TFrame f2;
        f2.c0 = f2.c1 = f2.c2 = f2.u0 = f2.u1 = f2.u2 = f2.u3 = f2.u4 = f2.u5 = f2.u6 = f2.u7 = f2.u8 = 0.0;
        if (valid) {
            const double *B_bb = R + 9;                         // B_bb[slot * 3 + c] for slots 1..6
            f2.c0 = R[0]; f2.c1 = R[1]; f2.c2 = R[2];
            f2.u0 = R[3]; f2.u1 = R[4]; f2.u2 = R[5]; f2.u3 = R[6]; f2.u4 = R[7]; f2.u5 = R[8];
            f2.u6 = R[9]; f2.u7 = R[10]; f2.u8 = R[11];
            if (cats & FR3D_CAT_SO) {
                if (s_so_possible(f1, par1, B_bb, bb2, s_r2[par1 + 1])) fl |= SCR_SO12;
                if (s_so_possible(f2, par2, A_bb, bb1, s_r2[par2 + 1])) fl |= SCR_SO21;
            }
            if ((cats & FR3D_CAT_SUGAR_RIBOSE) && par1 >= 0 && par2 >= 0 && par1 != 4 && par2 != 4 &&
                (bb1 >> 6 & 1u) && (bb2 >> 6 & 1u)) {
                const double o0 = A_bb[18] - B_bb[18], o1 = A_bb[19] - B_bb[19], o2 = A_bb[20] - B_bb[20];
                if (o0 * o0 + o1 * o1 + o2 * o2 < 14.4401) fl |= SCR_SR;
            }
            if ((cats & FR3D_CAT_BACKBONE) && par1 >= 0) {
                // some massive atom of a base-phosphate site of nt1 within 4.5 A of a backbone oxygen of nt2
                const int nsites = c_tab.n_sites[par1];
                bool any = false;
#pragma unroll 1
                for (int s = 0; s < nsites; ++s) {
                    const int ms = c_tab.site_m[par1][s], hs = c_tab.site_h[par1][s];
                    if (!((bm1 >> ms) & 1u) || !((bm1 >> hs) & 1u)) continue;        // NA:1937: both atoms needed
                    const double m0 = A_base[ms * 3], m1 = A_base[ms * 3 + 1], m2 = A_base[ms * 3 + 2];
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        const int oslot = (0x469213 >> (4 * i)) & 15;
                        const double *O = R + SCR_BB_OFF(oslot);
                        const double o0 = O[0], o1 = O[1], o2 = (oslot == 9 ? prev_o3_z : O[2]);
                        const double q0 = m0 - o0, q1 = m1 - o1, q2 = m2 - o2;
                        if (((bb2 >> oslot) & 1u) && (q0 * q0 + q1 * q1 + q2 * q2 < 20.2501)) {
                            // rare: every label also needs the angle massive-H-oxygen above 110 degrees; keep the
                            // pair unless it is certainly below 107 (cos >= 0 or sin^2 > 10.3 cos^2, tan^2 72.7 = 10.3)
                            const double h0 = A_base[hs * 3], h1 = A_base[hs * 3 + 1], h2 = A_base[hs * 3 + 2];
                            const double u0 = m0 - h0, u1 = m1 - h1, u2 = m2 - h2;
                            const double w0 = o0 - h0, w1 = o1 - h1, w2 = o2 - h2;
                            const double cs = u0 * w0 + u1 * w1 + u2 * w2;
                            const double x0 = u1 * w2 - u2 * w1, x1 = u2 * w0 - u0 * w2, x2 = u0 * w1 - u1 * w0;
                            if (!(cs >= 0.0 || x0 * x0 + x1 * x1 + x2 * x2 > 10.3 * cs * cs)) any = true;
                        }
                    }
                }
                if (any) fl |= SCR_BPH;
            }
        }
        // phase B: the base points of the 32 nt2 records take the place of the phase A rows
        __syncwarp();
#pragma unroll 4
        for (int k = 0; k < 32; ++k) {
            const int nb = __shfl_sync(FULL, n2, k);
            double *dst = rows + k * SCR_STRIDE + lane;
            cp_async8(dst, src_b0 + (size_t)nb * (FR3D_BASE_SLOTS * 3));
            if (lane < SCR_ROW_B - 32) cp_async8(dst + 32, src_b1 + (size_t)nb * (FR3D_BASE_SLOTS * 3));
        }
        cp_async_wait_all();
        __syncwarp();
        if (valid) {
            const double *B_base = R;
            if ((cats & FR3D_CAT_STACKING) && par1 >= 0 && par2 >= 0) {
                const unsigned c12 = s_cylinder_mask(f1, B_base, s_r2[FR3D_N_PARENTS + 2 + par1]) & bm2;
                const unsigned c21 = s_cylinder_mask(f2, A_base, s_r2[FR3D_N_PARENTS + 2 + par2]) & bm1;
                if (s_hull_any(f1, shull + par1 * 21, B_base, c12) || s_hull_any(f2, shull + par2 * 21, A_base, c21)) fl |= SCR_STACK;
            }
            if ((bm1 & 1u) && (bm2 & 1u) && par1 >= 0 && par2 >= 0) {
                unsigned ori_ok = 0;
                if (c_tab.combo_ok[par1 * FR3D_N_PARENTS + par2]) ori_ok |= 1u;
                if (c_tab.combo_ok[par2 * FR3D_N_PARENTS + par1]) ori_ok |= 2u;
                nZ = f1.u2 * f2.u2 + f1.u5 * f2.u5 + f1.u8 * f2.u8;
                if ((cats & FR3D_CAT_COPLANAR) && ori_ok && fabs(nZ) > 0.7757) {
                    const double e0 = f1.c0 - f2.c0, e1 = f1.c1 - f2.c1, e2 = f1.c2 - f2.c2;
                    const double lim = (0.3381 * 0.3381) * (e0 * e0 + e1 * e1 + e2 * e2);
                    const double d1 = e0 * f1.u2 + e1 * f1.u5 + e2 * f1.u8;
                    const double d2 = e0 * f2.u2 + e1 * f2.u5 + e2 * f2.u8;
                    if (d1 * d1 < lim && d2 * d2 < lim) fl |= SCR_CP;
                }
                const double gd0 = B_base[0] - A_base[0], gd1 = B_base[1] - A_base[1], gd2 = B_base[2] - A_base[2];
                const int sg = nZ > 0 ? 0 : 1;
                // stage 1, converged: displacement of each live orientation against the union box of its entry
                entry0 = (par1 * FR3D_N_PARENTS + par2) * 2 + sg;
                entry1 = (par2 * FR3D_N_PARENTS + par1) * 2 + sg;
#pragma unroll
                for (int ori = 0; ori < 2; ++ori) {
                    const TFrame &fa = ori ? f2 : f1;
                    const double s = ori ? -1.0 : 1.0;
                    const double v0 = s * gd0, v1 = s * gd1, v2 = s * gd2;
                    const double dZ = v0 * fa.u2 + v1 * fa.u5 + v2 * fa.u8;
                    const double dX = v0 * fa.u0 + v1 * fa.u3 + v2 * fa.u6;
                    const double dY = v0 * fa.u1 + v1 * fa.u4 + v2 * fa.u7;
                    if (ori) { dX1 = dX; dY1 = dY; dZ1 = dZ; } else { dX0 = dX; dY0 = dY; dZ0 = dZ; }
                    const double *ub = ubox + (ori ? entry1 : entry0) * 8;
                    const double cd = fmax(0.0, ub[0] - dX) + fmax(0.0, dX - ub[1]) + fmax(0.0, ub[2] - dY) +
                                      fmax(0.0, dY - ub[3]) + fmax(0.0, ub[4] - dZ) + fmax(0.0, dZ - ub[5]) +
                                      3.0 * (fmax(0.0, ub[6] - nZ) + fmax(0.0, nZ - ub[7]));
                    if (((ori_ok >> ori) & 1u) && !(fabs(dZ) > 3.6) && cd < 1.0 + 1e-9) {             // NA:2380
                        live |= 1u << ori;
                        // in-plane angle of this orientation (NA:2594), single precision is enough for the screen
                        const TFrame &fb = ori ? f1 : f2;
                        const float m11 = (float)(fa.u1 * fb.u1 + fa.u4 * fb.u4 + fa.u7 * fb.u7);
                        const float m01 = (float)(fa.u0 * fb.u1 + fa.u3 * fb.u4 + fa.u6 * fb.u7);
                        float ang = atan2f(m11, m01) * 57.29577951f - 90.f;
                        if (ang <= -90.f) ang += 360.f;
                        if (ori) ang1 = ang; else ang0 = ang;
                    }
                }
            }
        }
        // stage 2, warp cooperative: lane b evaluates box b of the entry of one live (pair, orientation) per step:
        // some box within 1.0 up to and including the angle term (NA:2546-2611)?  The gap term only adds.
        unsigned pending = __ballot_sync(FULL, live != 0);
        while (pending) {
            // task of this lane's half warp (two_tasks) or of the whole warp
            const int src_a = __ffs(pending) - 1;
            const unsigned rest = pending & (pending - 1);
            const int src_b = (two_tasks && rest) ? __ffs(rest) - 1 : -1;
            const bool second = two_tasks && (lane >> 4);
            const int srcl = second ? src_b : src_a;
            const int sub = two_tasks ? (lane & 15) : lane;
            const bool first_ori = live & 1u;
            const int from = srcl < 0 ? 0 : srcl;
            const float tX = __shfl_sync(FULL, (float)(first_ori ? dX0 : dX1), from);
            const float tY = __shfl_sync(FULL, (float)(first_ori ? dY0 : dY1), from);
            const float tZ = __shfl_sync(FULL, (float)(first_ori ? dZ0 : dZ1), from);
            const float tN = __shfl_sync(FULL, (float)nZ, from);
            const float tA = __shfl_sync(FULL, first_ori ? ang0 : ang1, from);
            const int te = __shfl_sync(FULL, first_ori ? entry0 : entry1, from);
            const int bstart = sidx[te * 2], bcount = srcl < 0 ? 0 : sidx[te * 2 + 1];
            bool hit = false;
            if (sub < bcount) {
                float cd, amin, amax;
                if (boxes_in_smem) {
                    const float *bx = sbox + (bstart + sub) * SCR_BOX_STRIDE;
                    cd = fmaxf(0.f, fabsf(tX - bx[0]) - bx[1]) + fmaxf(0.f, fabsf(tY - bx[2]) - bx[3]) +
                         fmaxf(0.f, fabsf(tZ - bx[4]) - bx[5]) + 3.f * fmaxf(0.f, fabsf(tN - bx[6]) - bx[7]);
                    amin = bx[8]; amax = bx[9];
                } else {
                    const double *bx = reinterpret_cast<const double *>(boxes + bstart + sub);
                    cd = fmaxf(0.f, (float)bx[0] - tX) + fmaxf(0.f, tX - (float)bx[1]) + fmaxf(0.f, (float)bx[2] - tY) +
                         fmaxf(0.f, tY - (float)bx[3]) + fmaxf(0.f, (float)bx[4] - tZ) + fmaxf(0.f, tZ - (float)bx[5]) +
                         3.f * (fmaxf(0.f, (float)bx[6] - tN) + fmaxf(0.f, tN - (float)bx[7]));
                    amin = (float)bx[8]; amax = (float)bx[9];
                }
                // the angle term (NA:2596-2611; ranges with amin > amax wrap around)
                const float below = fmaxf(0.f, amin - tA), above = fmaxf(0.f, tA - amax);
                cd += 0.05f * (amin < amax ? below + above : fminf(below, above));
                hit = cd < 1.0f + 1e-3f;
            }
            const unsigned hits = __ballot_sync(FULL, hit);
            if (lane == src_a || lane == src_b) {
                const bool any = two_tasks ? ((lane == src_b ? hits >> 16 : hits & 0xFFFFu) != 0) : hits != 0;
                if (any) { fl |= SCR_BP; live = 0; }
                else live &= live - 1;
            }
            pending = __ballot_sync(FULL, live != 0);
        }
        // warp-aggregated appends to the work lists: counters[4] light detail (sO, sugar-ribose), [5] stacking,
        // [6] coplanar / basepair, [7] base-phosphate
        const unsigned want_light = fl & (SCR_SO12 | SCR_SO21 | SCR_SR), want_stack = fl & SCR_STACK,
                       want_pair = fl & (SCR_CP | SCR_BP), want_bph = fl & SCR_BPH;
        const unsigned bl = __ballot_sync(FULL, want_light != 0);
        const unsigned bs = __ballot_sync(FULL, want_stack != 0);
        const unsigned bp = __ballot_sync(FULL, want_pair != 0);
        const unsigned bb = __ballot_sync(FULL, want_bph != 0);
        long long base = 0;
        if (lane == 0 && bl) base = (long long)atomicAdd((unsigned long long *)&counters[4], (unsigned long long)__popc(bl));
        if (lane == 1 && bs) base = (long long)atomicAdd((unsigned long long *)&counters[5], (unsigned long long)__popc(bs));
        if (lane == 2 && bp) base = (long long)atomicAdd((unsigned long long *)&counters[6], (unsigned long long)__popc(bp));
        if (lane == 3 && bb) base = (long long)atomicAdd((unsigned long long *)&counters[7], (unsigned long long)__popc(bb));
        const long long base_l = __shfl_sync(FULL, base, 0), base_s = __shfl_sync(FULL, base, 1),
                        base_p = __shfl_sync(FULL, base, 2), base_b = __shfl_sync(FULL, base, 3);
        const unsigned below = (1u << lane) - 1u;
        if (want_light) wl_light[base_l + __popc(bl & below)] = (int32_t)p;
        if (want_stack) wl_stack[base_s + __popc(bs & below)] = (int32_t)p;
        if (want_pair) wl_pair[base_p + __popc(bp & below)] = (int32_t)p;
        if (want_bph) wl_bph[base_b + __popc(bb & below)] = (int32_t)p;
    }
}

}  // namespace fr3d
