// K4, stacking and pairing work lists: thread per listed pair, records staged in TWO phases.
//
// classify_pairs_tpp_kernel<60, ...> keeps frame + base atoms of both nts of a warp's 32 pairs in shared memory
// (2 x 60 doubles per pair), which lets only 7 warps live on an SM, and ncu shows these kernels bound by fp64
// dependency latency at that occupancy (issue slots 2/3 idle).  Nothing in the stacking / coplanar / basepair
// rules needs both bases at once except the minimum atom-atom distance, and that loop keeps one base in registers
// anyway.  So the work is ordered by the base it reads and the row buffer is used twice:
//   phase A  row = frame 1 (12) | frame 2 (12) | glycosidic atom of nt1 (3) | base points of nt2 (48)
//            -> overlap of nt2 over nt1, cutoff boxes pass 1, gap12, nt2's points into registers
//   phase B  the base points of nt1 (48) overwrite those of nt2; the frames stay where they are
//            -> overlap of nt1 over nt2, gap21, minimum distance, labels, cutoff boxes pass 2, arbitration
// 75 doubles per pair instead of 121.  Registers are per SM sub-partition (16 K each), so a block can hold 4 x
// floor(16384 / (32 x regs)) warps: the stacking kernel fits 168 registers and runs 11 warps per SM (3 per
// sub-partition where needed), the pairing kernel needs ~200 (nt2's 48 coordinates stay in registers across the
// second staging) and runs 8.  The arithmetic is that of classify_tpp.cuh (same helper
// functions, same order of floating-point operations), so results are bit-identical to that kernel.
#pragma once

#include "classify_tpp.cuh"

namespace fr3d {

constexpr int L2_ROW = 75;               // doubles per pair; odd: lane k reads row k without bank conflicts
#ifndef FR3D_L2_WARPS_PAIR
#define FR3D_L2_WARPS_PAIR 8
#endif
constexpr int L2_WARPS_STACK = 11, L2_WARPS_PAIR = FR3D_L2_WARPS_PAIR;
// The pairing kernel also keeps the cutoff boxes in shared memory when there are at most L2_MAX_BOXES of them (the
// shipped tables have 243): 16 doubles per box at a stride of 17 doubles, so that lanes reading the same field of
// different boxes hit different banks.
constexpr int L2_MAX_BOXES = 256;
constexpr int L2_BOX_STRIDE = 17;
template <bool PAIRING, int WARPS>
__host__ __device__ constexpr size_t l2_smem_bytes() {
    return sizeof(double) * ((size_t)WARPS * 32 * L2_ROW + (PAIRING ? L2_MAX_BOXES * L2_BOX_STRIDE : 0));
}

// the 16 points of a base into registers; an absent atom sits at x = 1e150 so it can never be the minimum
__device__ __forceinline__ void t_mind_load(const double *base, uint32_t bm, double (&q)[FR3D_BASE_SLOTS * 3]) {
#pragma unroll
    for (int k = 0; k < FR3D_BASE_SLOTS; ++k) {
        const bool have = (bm >> k) & 1u;
        q[k * 3] = have ? base[k * 3] : 1e150;
        q[k * 3 + 1] = base[k * 3 + 1];
        q[k * 3 + 2] = base[k * 3 + 2];
    }
}

// calculate_base_min_distances (NA:1799-1829) of the points of base1 against the register-held points q of base2
// (same operations, in the same order, as the untwinned branch of t_mind)
__device__ __forceinline__ double t_mind_scan(const double *base1, uint32_t bm1, const double (&q)[FR3D_BASE_SLOTS * 3]) {
    double b0 = 1e300, b1 = 1e300;
#pragma unroll 1
    for (int v = 0; v < FR3D_BASE_SLOTS; ++v) {
        if (!((bm1 >> v) & 1u)) continue;
        const double a0 = base1[v * 3], a1 = base1[v * 3 + 1], a2 = base1[v * 3 + 2];
#pragma unroll
        for (int k = 0; k < FR3D_BASE_SLOTS; k += 2) {
            const double d0 = q[k * 3] - a0, d1 = q[k * 3 + 1] - a1, d2 = q[k * 3 + 2] - a2;
            const double dd = d0 * d0 + d1 * d1 + d2 * d2;
            if (dd < b0) b0 = dd;
            const double e0 = q[k * 3 + 3] - a0, e1 = q[k * 3 + 4] - a1, e2 = q[k * 3 + 5] - a2;
            const double ee = e0 * e0 + e1 * e1 + e2 * e2;
            if (ee < b1) b1 = ee;
        }
    }
    const double best = b1 < b0 ? b1 : b0;
    return best < 1e299 ? sqrt(best) : 9999.0;
}

// PAIRING = false: stacking list (base stacking only); true: pairing list (coplanar if in cats, basepair always)
template <bool PAIRING, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
classify_list2_kernel(fr3d_nts_t nts, const fr3d_box_t *__restrict__ boxes, const int32_t *__restrict__ box_index,
                      const int32_t *__restrict__ pair_i, const int32_t *__restrict__ pair_j,
                      const int32_t *__restrict__ pair_rank, int64_t pair_cap, uint32_t cats, int32_t index_offset,
                      fr3d_hit_t *__restrict__ hits, int64_t hit_cap, int64_t *counters,
                      const int32_t *__restrict__ worklist, int list_counter, const int64_t *list_len = nullptr) {
    extern __shared__ double l2_smem[];
    __shared__ double s_hull[FR3D_N_PARENTS * 7 * 3];        // a per-lane parent index would serialise constant loads
    __shared__ double s_r2in[FR3D_N_PARENTS];
    for (int item = threadIdx.x; item < FR3D_N_PARENTS * 7 * 3; item += blockDim.x)
        s_hull[item] = (&c_tab.hull[0][0][0])[item];
    if (threadIdx.x < FR3D_N_PARENTS) s_r2in[threadIdx.x] = c_tab.hull_r2in[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double *rows = l2_smem + (size_t)(threadIdx.x >> 5) * 32 * L2_ROW;
    // cutoff boxes: shared-memory copy (pairing kernel, small tables) or the caller's array
    const char *box_base = reinterpret_cast<const char *>(boxes);
    size_t box_stride = sizeof(fr3d_box_t);
    if (PAIRING) {
        __shared__ int s_nboxes;
        if (threadIdx.x == 0) s_nboxes = 0;
        __syncthreads();
        if (threadIdx.x < FR3D_N_PARENTS * FR3D_N_PARENTS * 2) {
            const int bstart = box_index[threadIdx.x * 2], bcount = box_index[threadIdx.x * 2 + 1];
            if (bcount > 0) atomicMax(&s_nboxes, bstart + bcount);
        }
        __syncthreads();
        if (s_nboxes <= L2_MAX_BOXES) {
            double *sbox = l2_smem + (size_t)WARPS * 32 * L2_ROW;
            for (int item = threadIdx.x; item < s_nboxes * 16; item += blockDim.x)
                sbox[(item >> 4) * L2_BOX_STRIDE + (item & 15)] = reinterpret_cast<const double *>(boxes + (item >> 4))[item & 15];
            box_base = reinterpret_cast<const char *>(sbox);
            box_stride = sizeof(double) * L2_BOX_STRIDE;
        }
        __syncthreads();
    }
    // per-lane sources of the three copies of a phase A row (item t = j * 32 + lane) and the two of a phase B row
    const double *src_a[3];
    int mult_a[3];
    bool first_a[3];                                         // item comes from nt1 (else nt2)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int t = j * 32 + lane;
        if (t < 3) { src_a[j] = nts.center + t; mult_a[j] = 3; first_a[j] = true; }
        else if (t < 12) { src_a[j] = nts.rot + (t - 3); mult_a[j] = 9; first_a[j] = true; }
        else if (t < 15) { src_a[j] = nts.center + (t - 12); mult_a[j] = 3; first_a[j] = false; }
        else if (t < 24) { src_a[j] = nts.rot + (t - 15); mult_a[j] = 9; first_a[j] = false; }
        else if (t < 27) { src_a[j] = nts.base_xyz + (t - 24); mult_a[j] = FR3D_BASE_SLOTS * 3; first_a[j] = true; }
        else { src_a[j] = nts.base_xyz + (t - 27); mult_a[j] = FR3D_BASE_SLOTS * 3; first_a[j] = false; }
    }
    const double *src_b0 = nts.base_xyz + lane, *src_b1 = nts.base_xyz + 32 + lane;

    long long n_pairs = list_len ? *list_len : counters[list_counter];       // (list_len: a length kept outside counters)
    if (n_pairs > pair_cap) n_pairs = pair_cap;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long rounds = (n_pairs + stride - 1) / stride;

    // the indices of a round are fetched one round ahead (list entry -> pair -> nt indices is a chain of
    // dependent global loads that would otherwise sit in front of every staging step)
    int nx1 = 0, nx2 = 0, nxr = 0;
    if (first < n_pairs) {
        const long long q = (long long)worklist[first];
        nx1 = pair_i[q]; nx2 = pair_j[q]; nxr = pair_rank[q];
    }
#pragma unroll 1
    for (long long r = 0; r < rounds; ++r) {
        const long long p = first + r * stride;
        const bool active = p < n_pairs;
        const int n1 = nx1, n2 = nx2, rank = nxr;
        nx1 = 0; nx2 = 0; nxr = 0;
        if (p + stride < n_pairs) {
            const long long q = (long long)worklist[p + stride];
            nx1 = pair_i[q]; nx2 = pair_j[q]; nxr = pair_rank[q];
        }
        if (!__any_sync(FULL, active)) continue;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(nts.base_mask + nx1));     // next round's masks
        asm volatile("prefetch.global.L1 [%0];" ::"l"(nts.base_mask + nx2));
        // ---------------- phase A staging
        __syncwarp();                                       // everybody is done with the previous round's rows
#pragma unroll 2
        for (int k = 0; k < 32; ++k) {
            const int na = __shfl_sync(FULL, n1, k), nb = __shfl_sync(FULL, n2, k);
            double *dst = rows + k * L2_ROW + lane;
            cp_async8(dst, src_a[0] + (size_t)(first_a[0] ? na : nb) * mult_a[0]);
            cp_async8(dst + 32, src_a[1] + (size_t)(first_a[1] ? na : nb) * mult_a[1]);
            if (lane + 64 < L2_ROW) cp_async8(dst + 64, src_a[2] + (size_t)(first_a[2] ? na : nb) * mult_a[2]);
        }
        uint32_t bm1 = 0, bm2 = 0;
        if (active) { bm1 = nts.base_mask[n1]; bm2 = nts.base_mask[n2]; }
        cp_async_wait_all();
        __syncwarp();

        // state carried from phase A to phase B
        const double *R = rows + lane * L2_ROW;
        double q2[FR3D_BASE_SLOTS * 3];
        double nZ = 0.0, z_2on1 = 0.0, gap12 = 100.0;
        bool nt2on1 = false, cp_frames = false, have_q2 = false;
        unsigned ori_ok = 0, bp_live = 0;
        uint32_t tw1 = 0, tw2 = 0;
        unsigned bx_live[2] = {0u, 0u};                     // surviving boxes, angle and displacement per orientation
        int bx_start[2] = {0, 0};
        double bx_ang[2] = {0.0, 0.0}, bx_d[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
        const int par1 = (int)((bm1 >> 16) & 15u) - 1, par2 = (int)((bm2 >> 16) & 15u) - 1;
        if (active) {
            const TFrame f1 = load_tframe_row(R), f2 = load_tframe_row(R + 12);
            const double *gly1 = R + 24, *base2 = R + 27;
            nZ = f1.u2 * f2.u2 + f1.u5 * f2.u5 + f1.u8 * f2.u8;                    // M[2][2] in both orientations
            if ((bm1 >> 22) & 1u) tw1 = (uint32_t)nts.meta[(size_t)n1 * FR3D_META_INTS + 7] >> 16;
            if ((bm2 >> 22) & 1u) tw2 = (uint32_t)nts.meta[(size_t)n2 * FR3D_META_INTS + 7] >> 16;
            bool want_q2 = false;
            if (!PAIRING) {
                // ---- base-base stacking, nt2 over nt1 (NA:1602-1702)
                if (par1 >= 0 && par2 >= 0) {
                    nt2on1 = t_overlap(f1, s_hull + par1 * 21, s_r2in[par1], base2, bm2, z_2on1);
                    want_q2 = true;
                }
            } else {
                // ---- coplanar + basepair, the parts that read nt2's base (NA:834-1013)
                const bool gly_ok = (bm1 & 1u) && (bm2 & 1u) && par1 >= 0 && par2 >= 0;
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
                    const double gd0 = base2[0] - gly1[0], gd1 = base2[1] - gly1[1], gd2 = base2[2] - gly1[2];
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
                        bx_live[ori] = t_boxes_live(box_base, box_stride, bstart, bcount, dX, dY, dZ, nZ, M11, M01, bx_ang[ori]);
                        bx_start[ori] = bstart;
                        bx_d[ori][0] = dX; bx_d[ori][1] = dY; bx_d[ori][2] = dZ;
                        if (bx_live[ori]) bp_live |= 1u << ori;
                    }
                }
                if (bp_live || cp_frames) {
                    gap12 = t_gap_inline(f1, f2, par2, base2, bm2, tw2);
                    want_q2 = true;
                }
            }
            if (want_q2 && !tw1 && !tw2) {
                t_mind_load(base2, bm2, q2);
                have_q2 = true;
            }
        }
        // ---------------- phase B staging: the base points of nt1 replace the phase A rows
        __syncwarp();
#pragma unroll 4
        for (int k = 0; k < 32; ++k) {
            const int na = __shfl_sync(FULL, n1, k);
            double *dst = rows + k * L2_ROW + 27 + lane;
            cp_async8(dst, src_b0 + (size_t)na * (FR3D_BASE_SLOTS * 3));
            if (lane < FR3D_BASE_SLOTS * 3 - 32) cp_async8(dst + 32, src_b1 + (size_t)na * (FR3D_BASE_SLOTS * 3));
        }
        cp_async_wait_all();
        __syncwarp();

        unsigned hmask = 0;
        int c_stack = 0, c_bp = 0, bp_box = 0;
        double bp_cd = 0.0, bp_gap = 0.0;
        if (active) {
            const TFrame f1 = load_tframe_row(R), f2 = load_tframe_row(R + 12);    // (the frames were left in place)
            const double *base1 = R + 27;
            // twins (modified residues with ideal-position copies, rare): both bases at once, nt2's from global memory
            const double *base2_g = nts.base_xyz + (size_t)n2 * FR3D_BASE_SLOTS * 3;
            double mind = 9999.0;
            if (!PAIRING) {
                int stack_code = -1;
                bool stack_both = false;
                if (par1 >= 0 && par2 >= 0) {
                    double z_1on2;
                    const bool nt1on2 = t_overlap(f2, s_hull + par2 * 21, s_r2in[par2], base1, bm1, z_1on2);
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
                if (stack_code >= 0) {
                    mind = have_q2 ? t_mind_scan(base1, bm1, q2)
                                   : t_mind(f1, f2, par1, par2, base1, base2_g, bm1, bm2, tw1, tw2);
                    if (!(fabs(mind) < 1.0)) {                                    // NA:1707-1718
                        int near = 1;
                        if (fabs(mind) < 4.0 && fabs(mind) > 1.0 && fabs(nZ) > 0.6 && stack_both) near = 0;
                        hmask |= 1u << FR3D_SLOT_STACK;
                        c_stack = stack_code | (near << 2);
                    }
                }
            } else if (bp_live || cp_frames) {
                const double gap21 = t_gap_inline(f2, f1, par1, base1, bm1, tw1);
                const bool cp_gap = cp_frames && (((ori_ok & 1u) && gap12 < 1.5179) || ((ori_ok & 2u) && gap21 < 1.5179));
                if (bp_live || cp_gap)
                    mind = have_q2 ? t_mind_scan(base1, bm1, q2)
                                   : t_mind(f1, f2, par1, par2, base1, base2_g, bm1, bm2, tw1, tw2);
                if (cp_gap) {                                                     // NA:2080-2096
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
                        t_boxes_final(box_base, box_stride, bx_start[ori], bx_live[ori], bx_d[ori][0], bx_d[ori][1], bx_d[ori][2], nZ,
                                      bx_ang[ori], gmax, mind, ori ? par1 : par2, res[ori]);
                    }
                    int use = -1;                                                 // NA:935-957
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
                    const bool isbp = s == FR3D_SLOT_BP;
                    const int code = s == FR3D_SLOT_STACK ? c_stack : (isbp ? c_bp : 0);
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
