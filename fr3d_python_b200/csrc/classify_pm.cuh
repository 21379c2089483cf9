// K4, phase-major form: each warp takes PM_K consecutive candidate pairs, stages all of them in
// shared memory with cp.async (deep memory-level parallelism, no register staging), and then runs
// every rule family as a short loop over the PM_K staged pairs.
//
// Why: ncu on the pair-major kernel (classify.cuh) showed the warps stalled on instruction fetch
// (stall_no_instruction at every 128-byte line of a ~70 KB straight-line body that is executed once
// per pair and never re-used out of the L0 instruction cache) and on the dependent index -> record
// loads.  Here each phase body is executed PM_K times back to back, and the loads of PM_K pairs are
// in flight together.  The rules themselves are the same code as in classify.cuh (same helpers, same
// arithmetic, same order of floating-point operations).
#pragma once

#include <cstddef>

#include "classify.cuh"

namespace fr3d {

static constexpr int PM_K = 4;
// internal category bit: skip the coplanar / basepair phases (the kernel then serves as the fp64 fallback of a single
// rule family, e.g. the undecided pairs of the single-precision stacking kernel)
#define PM_CAT_NO_PAIRING 0x80000000u
#ifndef PM_UNROLL
#define PM_UNROLL 1
#endif
static constexpr int PM_UNROLL_N = PM_UNROLL;   // pairs of a phase loop interleaved by the compiler

struct alignas(16) PairState {
    WarpScratch d;
    double cd[2][32];        // per lane: running cutoff_distance of "my" box, per orientation
    double gap12, gap21, mind;
    double bp_cd;
    int n1, n2, rank;
    uint32_t bm1, bm2, bb1, bb2;
    unsigned hmask;
    unsigned flags;          // bit0 cp_frames, bits1-2 ori_ok, bits3-4 bp_live, bit5 stack_both
    int stack_code;
    int bp_box;
    int code[9];
};

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem_src) : "memory");
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem_src) : "memory");
}

__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// ---- bulk asynchronous copies (TMA, non-tensor form) completing on an mbarrier: one instruction moves a whole record
// row (a multiple of 16 bytes, 16-byte aligned on both sides) instead of one cp.async per 16-byte piece and lane.
__device__ __forceinline__ uint32_t smem_addr_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(arrivals) : "memory");
}

// orders generic-proxy accesses to shared memory (the mbarrier init, the lanes' reads of a row buffer) before later
// async-proxy ones (the bulk copy engine's writes into the same buffer)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_addr_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_wait_parity(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra MBAR_WAIT_%=;\n"
        "}\n" ::"r"(smem_addr_u32(bar)),
        "r"(parity)
        : "memory");
}

#define PM_C1(k) (S.fr[(k)])
#define PM_U1(k) (S.fr[3 + (k)])
#define PM_C2(k) (S.fr[12 + (k)])
#define PM_U2(k) (S.fr[15 + (k)])

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, FR3D_CLASSIFY_MIN_BLOCKS)
classify_pairs_pm_kernel(fr3d_nts_t nts, const fr3d_box_t *__restrict__ boxes, const int32_t *__restrict__ box_index,
                         const int32_t *__restrict__ pair_i, const int32_t *__restrict__ pair_j,
                         const int32_t *__restrict__ pair_rank, int64_t pair_cap, uint32_t cats, int32_t index_offset,
                         fr3d_hit_t *__restrict__ hits, int64_t hit_cap, int64_t *counters,
                         const int32_t *__restrict__ worklist, const int64_t *list_len = nullptr) {
    __shared__ PairState s_state[WARPS_PER_BLOCK][PM_K];

    const int lane = threadIdx.x & 31;
    PairState *W = s_state[threadIdx.x >> 5];
    const long long warp0 = (long long)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * WARPS_PER_BLOCK;
    // with a work list (classify_screen.cuh) the kernel walks its entries instead of the whole pair list
    // (list_len: the length of a work list kept outside counters - the undecided pairs of classify_stack32.cuh)
    long long n_pairs = list_len ? *list_len : (worklist ? counters[5] : counters[0]);
    if (n_pairs > pair_cap) n_pairs = pair_cap;
    const int half = lane >> 4;      // 0: atoms of nt2 seen from frame 1, 1: atoms of nt1 seen from frame 2
    const int a = lane & 15;

    // ---- per-lane copy plan (loop invariant).  A nt's base block is 384 contiguous bytes (24 x 16 B), its
    // backbone block 240 bytes (15 x 16 B); both are 16-byte aligned.  Shared layout: atoms[0..15] = nt2,
    // atoms[16..31] = nt1, bb[0..9] = nt2, bb[10..19] = nt1, fr[0..11] = nt1 frame, fr[12..23] = nt2 frame.
    const char *base_b = reinterpret_cast<const char *>(nts.base_xyz);
    const char *bb_b = reinterpret_cast<const char *>(nts.bb_xyz);
    const int off_atoms = (int)offsetof(PairState, d) + (int)offsetof(WarpScratch, atoms);
    const int off_bb = (int)offsetof(PairState, d) + (int)offsetof(WarpScratch, bb);
    const int off_fr = (int)offsetof(PairState, d) + (int)offsetof(WarpScratch, fr);
    // copy 1 (16 B): lanes 0-23 -> nt2 base chunks 0-23, lanes 24-31 -> nt1 base chunks 0-7
    const bool c1_n1 = lane >= 24;
    const char *c1_src = base_b + (lane < 24 ? lane : lane - 24) * 16;
    const int c1_dst = off_atoms + (lane < 24 ? lane * 16 : 384 + (lane - 24) * 16);
    // copy 2 (16 B): lanes 0-15 -> nt1 base chunks 8-23, lanes 16-30 -> nt2 backbone chunks 0-14
    const bool c2_n1 = lane < 16;
    const char *c2_src = lane < 16 ? base_b + (8 + lane) * 16 : bb_b + (lane - 16) * 16;
    const size_t c2_stride = lane < 16 ? (size_t)FR3D_BASE_SLOTS * 24 : (size_t)FR3D_BB_SLOTS * 24;
    const int c2_dst = lane < 16 ? off_atoms + 384 + (8 + lane) * 16 : off_bb + (lane - 16) * 16;
    // copy 3 (16 B): lanes 0-14 -> nt1 backbone chunks
    const char *c3_src = bb_b + lane * 16;
    const int c3_dst = off_bb + 240 + lane * 16;
    // copy 4 (8 B): lanes 0-11 -> nt1 centre (3) + rotation (9), lanes 12-23 -> nt2
    const int fe = lane < 12 ? lane : lane - 12;
    const char *c4_src = fe < 3 ? reinterpret_cast<const char *>(nts.center) + fe * 8
                                : reinterpret_cast<const char *>(nts.rot) + (fe - 3) * 8;
    const size_t c4_stride = fe < 3 ? 24 : 72;
    const int c4_dst = off_fr + lane * 8;

    // pairs per warp and round: PM_K, fewer when the list is short (the fp64 fallback of the stacking kernel hands
    // over ~2 k pairs: one pair per warp on 1.8 k warps is a quarter of the latency of four pairs on 450)
    const long long per_warp = (n_pairs + nwarps - 1) / nwarps;
    const int kk = per_warp < 1 ? 1 : (per_warp < PM_K ? (int)per_warp : PM_K);
#pragma unroll 1
    for (long long p0 = warp0 * kk; p0 < n_pairs; p0 += nwarps * kk) {
        const int nk = (int)((n_pairs - p0) < kk ? (n_pairs - p0) : kk);

        // ---------------- phase 0: indices, then every record of the PM_K pairs with cp.async
        __syncwarp();
        if (lane < nk) {
            const long long q = worklist ? (long long)worklist[p0 + lane] : p0 + lane;
            W[lane].n1 = pair_i[q];
            W[lane].n2 = pair_j[q];
            W[lane].rank = pair_rank[q];
        }
        __syncwarp();
#pragma unroll 1
        for (int k = 0; k < nk; ++k) {
            PairState &P = W[k];
            const int n1 = P.n1, n2 = P.n2;
            // five copy instructions per pair; every lane's (array, nt, offset, destination) is loop invariant
            char *sb = reinterpret_cast<char *>(&P);
            cp_async16(sb + c1_dst, c1_src + (size_t)(c1_n1 ? n1 : n2) * (FR3D_BASE_SLOTS * 24));
            if (lane < 31) cp_async16(sb + c2_dst, c2_src + (size_t)(c2_n1 ? n1 : n2) * c2_stride);
            if (lane < 15) cp_async16(sb + c3_dst, c3_src + (size_t)n1 * (FR3D_BB_SLOTS * 24));
            if (lane < 24) cp_async8(sb + c4_dst, c4_src + (size_t)(lane < 12 ? n1 : n2) * c4_stride);
            if (lane < 4) cp_async4(&P.bm1 + lane, (lane < 2 ? nts.base_mask : nts.bb_mask) + ((lane & 1) ? n2 : n1));
            if (lane == 4) { P.hmask = 0; P.flags = 0; P.stack_code = -1; }
        }
        cp_async_wait_all();
        __syncwarp();

        // ---------------- phase A: base-oxygen stacking, both directions (NA:758-778)
        if (cats & FR3D_CAT_SO) {
#pragma unroll PM_UNROLL_N
            for (int k = 0; k < nk; ++k) {
                PairState &P = W[k];
                const WarpScratch &S = P.d;
                const uint32_t bb1 = P.bb1, bb2 = P.bb2;
                const int par1 = (int)((P.bm1 >> 16) & 15u) - 1, par2 = (int)((P.bm2 >> 16) & 15u) - 1;
                const int g = (lane >> 3) & 1, o = lane & 7;
                const bool act = lane < 16 && o < 6;
                const int pb = g ? par2 : par1;
                const int oslot = (0x213456 >> (4 * o)) & 7;         // O2' O3' O4' O5' OP1 OP2 -> slots 6 5 4 3 1 2
                const bool present = act && (((g ? bb1 : bb2) >> oslot) & 1u);
                double ox, oy, oz;
                const double *ob = S.bb[(g ? 10 : 0) + (act ? oslot : 0)];
                project(g ? &S.fr[12] : &S.fr[0], ob[0], ob[1], ob[2], ox, oy, oz);
                const bool has_tab = pb >= 0 && pb < FR3D_N_PARENTS;
                const bool maybe = present && has_tab && fabs(oz) < 3.6 && (ox * ox + oy * oy) < 25.0;
                if (__any_sync(FULL, maybe)) {
                    bool ring = false, nr = false;
                    if (maybe) {
                        const bool has_div = c_tab.has_div[pb];
                        const bool left = has_div && left_of(c_tab.ring_div[pb], ox, oy);
                        if (!(fabs(oz) < 2.0)) {
                            ring = true;
                            if (left) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) ring = ring && left_of(c_tab.ring5[pb][q], ox, oy);
                            } else {
#pragma unroll
                                for (int q = 0; q < 6; ++q) ring = ring && left_of(c_tab.ring6[pb][q], ox, oy);
                            }
                        }
                        if (fabs(oz) < 3.5) nr = in_ellipse(left ? c_tab.ell5[pb] : c_tab.ell6[pb], ox, oy);
                    }
                    const unsigned ring_b = __ballot_sync(FULL, ring);
                    const unsigned true_b = __ballot_sync(FULL, ring && fabs(oz) < 3.5);
                    const unsigned gmask = 0xFFu << (g * 8);
                    const bool use_ring = (ring_b & gmask) != 0;
                    double key = use_ring ? (ring ? fabs(oz) : 1e300) : (nr ? (ox * ox + oy * oy) : 1e300);
                    int idx = o;
                    group_argmin<8>(key, idx);
                    const double zsel = __shfl_sync(FULL, oz, (lane & ~7) | idx);
                    int code = -1;
                    if (key < 1e299) code = idx | ((zsel > 0) ? 0 : 8) | ((use_ring && (true_b & gmask)) ? 0 : 16);
                    const int code12 = __shfl_sync(FULL, code, 0);
                    const int code21 = __shfl_sync(FULL, code, 8);
                    if (lane == 0) {
                        if (code12 >= 0) { P.hmask |= 1u << FR3D_SLOT_SO12; P.code[FR3D_SLOT_SO12] = code12; }
                        if (code21 >= 0) { P.hmask |= 1u << FR3D_SLOT_SO21; P.code[FR3D_SLOT_SO21] = code21; }
                    }
                }
            }
        }

        // ---------------- phase B: base-base stacking, part 1 (NA:1602-1702)
        if (cats & FR3D_CAT_STACKING) {
#pragma unroll PM_UNROLL_N
            for (int k = 0; k < nk; ++k) {
                PairState &P = W[k];
                const WarpScratch &S = P.d;
                const int par1 = (int)((P.bm1 >> 16) & 15u) - 1, par2 = (int)((P.bm2 >> 16) & 15u) - 1;
                if (par1 < 0 || par2 < 0) continue;
                const bool valid = ((half ? P.bm1 : P.bm2) >> a) & 1u;
                double x, y, z;
                project(half ? &S.fr[12] : &S.fr[0], S.atoms[lane][0], S.atoms[lane][1], S.atoms[lane][2], x, y, z);
                const bool inside = valid && in_hull(half ? par2 : par1, x, y, z);
                const unsigned in_b = __ballot_sync(FULL, inside);
                if (!in_b) continue;
                const double nZ = PM_U1(2) * PM_U2(2) + PM_U1(5) * PM_U2(5) + PM_U1(8) * PM_U2(8);
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
                    if (lane == 0) {
                        P.stack_code = code;
                        if (nt2on1 && nt1on2) P.flags |= 32u;
                    }
                }
            }
        }

        // ---------------- phase C: sugar-ribose, both directions (NA:791-804)
        if (cats & FR3D_CAT_SUGAR_RIBOSE) {
#pragma unroll 1
            for (int k = 0; k < nk; ++k) {
                PairState &P = W[k];
                const WarpScratch &S = P.d;
                const uint32_t bb1 = P.bb1, bb2 = P.bb2, bm1 = P.bm1, bm2 = P.bm2;
                const int par1 = (int)((bm1 >> 16) & 15u) - 1, par2 = (int)((bm2 >> 16) & 15u) - 1;
                if (!(par1 >= 0 && par2 >= 0 && par1 != 4 && par2 != 4 && (bb1 >> 6 & 1u) && (bb2 >> 6 & 1u))) continue;
                const double o0 = S.bb[16][0] - S.bb[6][0], o1 = S.bb[16][1] - S.bb[6][1], o2 = S.bb[16][2] - S.bb[6][2];
                const double oo = o0 * o0 + o1 * o1 + o2 * o2;
                if (oo < 14.4401 && !(sqrt(oo) > 3.8)) {
                    const int s1 = c_tab.sr_slot[par1], s2 = c_tab.sr_slot[par2];
                    const int r12 = sugar_ribose_tail(&S.bb[10], &S.bb[0], bb1, bb2, s1 >= 0 && ((bm1 >> s1) & 1u),
                                                      S.atoms[16 + (s1 < 0 ? 0 : s1)], PM_U1(2), PM_U1(5), PM_U1(8));
                    const int r21 = sugar_ribose_tail(&S.bb[0], &S.bb[10], bb2, bb1, s2 >= 0 && ((bm2 >> s2) & 1u),
                                                      S.atoms[s2 < 0 ? 0 : s2], PM_U2(2), PM_U2(5), PM_U2(8));
                    if (lane == 0) {
                        if (r12 >= 0) { P.hmask |= 1u << FR3D_SLOT_SR12; P.code[FR3D_SLOT_SR12] = r12; }
                        if (r21 >= 0) { P.hmask |= 1u << FR3D_SLOT_SR21; P.code[FR3D_SLOT_SR21] = r21; }
                    }
                }
            }
        }

        // ---------------- phase D: base-phosphate / base-ribose, nt1 base -> nt2 backbone (NA:807-829)
        if (cats & FR3D_CAT_BACKBONE) {
#pragma unroll 1
            for (int k = 0; k < nk; ++k) {
                PairState &P = W[k];
                WarpScratch &S = P.d;
                const uint32_t bb2 = P.bb2, bm1 = P.bm1;
                const int par1 = (int)((bm1 >> 16) & 15u) - 1;
                if (par1 < 0) continue;
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
                            cand = dst < (c_tab.site_carbon[par1][site] == 0 ? 16.0001 : 20.2501);
                        }
                    }
                }
                if (__any_sync(FULL, cand)) {
                    double ang = 0.0;
                    if (cand) { ang = hb_angle(S.atoms[16 + ms], S.atoms[16 + hs], S.bb[oslot]); dst = sqrt(dst); }
                    else dst = 0.0;
                    __syncwarp();
                    S.ang[lane] = ang; S.dst[lane] = dst;
                    __syncwarp();
                    if (lane == 0) {
                        const int sel = backbone_select(S, par1, bb2);
                        if ((sel & 0xFF) != 0xFF) { P.hmask |= 1u << FR3D_SLOT_BPH; P.code[FR3D_SLOT_BPH] = sel & 0xFF; }
                        if (((sel >> 8) & 0xFF) != 0xFF) { P.hmask |= 1u << FR3D_SLOT_BR; P.code[FR3D_SLOT_BR] = (sel >> 8) & 0xFF; }
                    }
                }
            }
        }

        // ---------------- phase E: coplanar frame terms + basepair cutoff boxes (part 1)
#pragma unroll 1
        for (int k = 0; k < nk; ++k) {
            PairState &P = W[k];
            const WarpScratch &S = P.d;
            const uint32_t bm1 = P.bm1, bm2 = P.bm2;
            const int par1 = (int)((bm1 >> 16) & 15u) - 1, par2 = (int)((bm2 >> 16) & 15u) - 1;
            const bool gly_ok = (bm1 & 1u) && (bm2 & 1u) && par1 >= 0 && par2 >= 0;   // NA:674, 834-837
            if (!gly_ok || (cats & PM_CAT_NO_PAIRING)) continue;
            unsigned ori_ok = 0, bp_live = 0;
            bool cp_frames = false;
            if (c_tab.combo_ok[par1 * FR3D_N_PARENTS + par2]) ori_ok |= 1u;
            if (c_tab.combo_ok[par2 * FR3D_N_PARENTS + par1]) ori_ok |= 2u;
            const double nZ = PM_U1(2) * PM_U2(2) + PM_U1(5) * PM_U2(5) + PM_U1(8) * PM_U2(8);
            if ((cats & FR3D_CAT_COPLANAR) && ori_ok && fabs(nZ) > 0.7757) {
                const double e0 = PM_C1(0) - PM_C2(0), e1 = PM_C1(1) - PM_C2(1), e2 = PM_C1(2) - PM_C2(2);
                const double lim = (0.3381 * 0.3381) * (e0 * e0 + e1 * e1 + e2 * e2);
                const double d1 = e0 * PM_U1(2) + e1 * PM_U1(5) + e2 * PM_U1(8);
                const double d2 = e0 * PM_U2(2) + e1 * PM_U2(5) + e2 * PM_U2(8);
                cp_frames = d1 * d1 < lim && d2 * d2 < lim;
            }
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
                P.cd[ori][lane] = ok ? cd : 1e300;
            }
            if (lane == 0) P.flags |= (cp_frames ? 1u : 0u) | (ori_ok << 1) | (bp_live << 3);
        }
        __syncwarp();

        // ---------------- phase F: gaps, min distance, final decisions, emission
#pragma unroll 1
        for (int k = 0; k < nk; ++k) {
            PairState &P = W[k];
            const WarpScratch &S = P.d;
            const uint32_t bm1 = P.bm1, bm2 = P.bm2;
            const unsigned flags = P.flags;
            const int stack_code = P.stack_code;
            unsigned hmask = P.hmask;
            const bool cp_frames = flags & 1u;
            const unsigned ori_ok = (flags >> 1) & 3u, bp_live = (flags >> 3) & 3u;
            int c_stack = 0, c_bp = 0, bp_box = 0;
            double bp_cd = 0.0;
            double gap12 = 100.0, gap21 = 100.0, mind = 9999.0;
            if (stack_code >= 0 || bp_live || cp_frames) {
                const int par1 = (int)((bm1 >> 16) & 15u) - 1, par2 = (int)((bm2 >> 16) & 15u) - 1;
                const double nZ = PM_U1(2) * PM_U2(2) + PM_U1(5) * PM_U2(5) + PM_U1(8) * PM_U2(8);
                const bool valid = ((half ? bm1 : bm2) >> a) & 1u;
                // modified residues with ideal-position copies: both points of a twin slot count (rare path)
                const bool twins = ((bm1 | bm2) >> 22) & 1u;
                uint32_t tw1 = 0, tw2 = 0;
                if (twins) {
                    if ((bm1 >> 22) & 1u) tw1 = (uint32_t)nts.meta[(size_t)P.n1 * FR3D_META_INTS + 7] >> 16;
                    if ((bm2 >> 22) & 1u) tw2 = (uint32_t)nts.meta[(size_t)P.n2 * FR3D_META_INTS + 7] >> 16;
                }
                if (twins && (bp_live || cp_frames)) {
                    gaps_twins(S, bm1, bm2, tw1, tw2, par1, par2, lane, gap12, gap21);
                } else if (bp_live || cp_frames) {            // calculate_basepair_gap, NA:2189-2259, both directions
                    const double *fo = half ? &S.fr[12] : &S.fr[0];
                    const double px = S.atoms[lane][0], py = S.atoms[lane][1], pz = S.atoms[lane][2];
                    const double w0 = px - fo[0], w1 = py - fo[1], w2 = pz - fo[2];
                    const double z = w0 * fo[5] + w1 * fo[8] + w2 * fo[11];
                    double key = valid ? (w0 * w0 + w1 * w1 + w2 * w2) : 1e300;
                    double g = 100.0;
#pragma unroll 1
                    for (int r = 0; r < 3; ++r) {
                        double kk = key; int idx = a;
                        group_argmin<16>(kk, idx);
                        const double zz = fabs(__shfl_sync(FULL, z, (lane & 16) | idx));
                        if (kk < 1e299) {
                            if (zz < g) g = zz;
                            if (a == idx) key = 1e300;
                        }
                    }
                    gap12 = __shfl_sync(FULL, g, 0);
                    gap21 = __shfl_sync(FULL, g, 16);
                }
                const bool cp_gap = cp_frames && (((ori_ok & 1u) && gap12 < 1.5179) || ((ori_ok & 2u) && gap21 < 1.5179));
                if (twins && (stack_code >= 0 || bp_live || cp_gap)) {
                    mind = mind_twins(S, bm1, bm2, tw1, tw2, par1, par2, lane);
                } else if (stack_code >= 0 || bp_live || cp_gap) {   // calculate_base_min_distances, NA:1799-1829
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
                // stacking, part 2 (NA:1707-1718)
                if (stack_code >= 0 && !(fabs(mind) < 1.0)) {
                    int near = 1;
                    if (fabs(mind) < 4.0 && fabs(mind) > 1.0 && fabs(nZ) > 0.6 && (flags & 32u)) near = 0;
                    hmask |= 1u << FR3D_SLOT_STACK;
                    c_stack = stack_code | (near << 2);
                }
                // coplanar, part 2 (NA:2080-2096)
                if (cp_gap) {
                    const bool both_std = (bm1 >> 20 & 1u) && (bm2 >> 20 & 1u);
                    if (mind < 3.4589 && !(mind >= 2.4589 && both_std)) hmask |= 1u << FR3D_SLOT_CP;
                }
                // basepair, part 2 (NA:2642-2759) and the 12-vs-21 arbitration (NA:935-957)
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
                        double cd = P.cd[ori][lane];
                        const bool ok = cd < 1e299;
                        const fr3d_box_t *bx = boxes + bstart + (ok ? lane : 0);
                        bool is_match = false, is_near = false;
                        int bflags = 0;
                        if (ok) {
                            bflags = bx->flags;
                            const double gapmax = bx->gapmax;
                            if (gapmax > 0.1) cd += 4.0 * fmax(0.0, gmax - gapmax);        // NA:2643-2645
                            bool one_hb = false;                                           // NA:2648-2652
                            if ((bflags & 2) && (pb == 1 || pb == 3)) one_hb = true;
                            else if ((bflags & 4) && pb == 3) one_hb = true;
                            const bool starts_n = bflags & 1;
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
                            double key = 1e300;
                            if (match_b) { if (is_match) key = (double)(bx->subcat * 64 + bx->name_len); }
                            else if (is_near) key = cd;
                            int idx = lane;
                            group_argmin<32>(key, idx);
                            const double cds = __shfl_sync(FULL, cd, idx);
                            const int fls = __shfl_sync(FULL, bflags, idx);
                            box_sel[ori] = bstart + idx;
                            cd_sel[ori] = cds;
                            near_sel[ori] = ((cds > 0 && !(fls & 1)) ? 1 : 0) | ((fls & 1) ? 2 : 0);
                        }
                    }
                    int use = -1;
                    if (box_sel[0] >= 0 && box_sel[1] >= 0) {
                        const bool n0 = near_sel[0] != 0, n1_ = near_sel[1] != 0;
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
            }
            // emit: lane s writes hit slot s, compacted
            if (hmask) {
                long long base = 0;
                if (lane == 0) base = (long long)atomicAdd((unsigned long long *)&counters[1], (unsigned long long)__popc(hmask));
                base = __shfl_sync(FULL, base, 0);
                if ((hmask >> lane) & 1u) {
                    const long long pos = base + __popc(hmask & ((1u << lane) - 1u));
                    if (pos < hit_cap) {
                        int mycode = 0;
                        if (lane == FR3D_SLOT_STACK) mycode = c_stack;
                        else if (lane == FR3D_SLOT_BP) mycode = c_bp;
                        else if (lane != FR3D_SLOT_CP) mycode = P.code[lane];
                        const bool isbp = lane == FR3D_SLOT_BP;
                        const double cdv = isbp ? bp_cd : 0.0, gv = isbp ? fmax(gap12, gap21) : 0.0;
                        const uint32_t w3 = (uint32_t)lane | ((uint32_t)mycode << 8) | ((uint32_t)(isbp ? bp_box : 0) << 16);
                        uint4 *dst = reinterpret_cast<uint4 *>(hits + pos);
                        dst[0] = make_uint4((uint32_t)(P.n1 + index_offset), (uint32_t)(P.n2 + index_offset),
                                            (uint32_t)(P.rank + 16 * index_offset), w3);
                        dst[1] = make_uint4((uint32_t)__double2loint(cdv), (uint32_t)__double2hiint(cdv),
                                            (uint32_t)__double2loint(gv), (uint32_t)__double2hiint(gv));
                    }
                }
            }
        }
    }
}

#undef PM_C1
#undef PM_U1
#undef PM_C2
#undef PM_U2

}  // namespace fr3d
