// K4a, single-precision form of the screen (classify_screen.cuh holds the rationale of the screen itself and the
// double-precision kernel this one replaced; FR3D_SCREEN=64 still selects that one for A/B runs).
//
// The screen only decides which pairs go on to the fp64 work-list kernels, and every test in it is a NECESSARY
// condition of a rule family.  So it needs no fp64: it runs on a compact single-precision copy of the nucleotide
// records with every threshold relaxed by a margin that is orders of magnitude above the fp32 rounding error
// (bounds below), which keeps it a superset.  ncu on the fp64 screen (profiles/step_r1n_*): 16 warps per SM (the
// 49-double rows fill shared memory, 128 registers per thread), 41 % issue utilisation, stalls = global loads of
// nt1 and fp64 dependency chains.  The fp32 record is 320 B instead of 760 B (the 316 k records of the bench batch
// are 101 MB: they mostly stay in the 126 MB L2 between the kernel that writes them and this one), a staged row is
// 52 floats (copied in 16-byte pieces, read with 128-bit loads), and 24 warps run per SM at 80 registers.
//
// Screen record of one nt, 80 32-bit words (screen_records_kernel):
//   0..2   base centre as fixed point, 2^-16 A (int32): e = c2 - c1 is an exact integer difference
//   3..5   first column of the rotation matrix (x axis of the base frame),  6..8  third column (the base normal);
//          the second column is their cross product (the fit always returns a proper rotation,
//          superpositions.py:71-78)
//   9..29  OP1 OP2 O5' O4' O3' O2' and the previous nt's O3' (backbone slots 1..6, 9) minus the centre
//   30     base_mask   31  bb_mask
//   32..79 the 16 base points minus the centre
// Words 0..31 are staged for the sO / sugar-ribose / base-phosphate screens (eight 16-byte copies per row: one
// instruction moves four rows), words 32..79 then take their place for the stacking / pairing screens.
//
// Error bounds (|coordinate| < 32 000 A is enforced by the fixed-point conversion, larger ones raise
// FR3D_E_RANGE): centre-relative coordinates are below 16 A, so their fp32 images carry <= 1e-6 A; e carries
// <= 1.6e-5 A; a projection on a frame axis adds <= 1e-5 A; a squared length below 600 A^2 carries <= 2e-4 A^2.
// Margins: 1e-3 A on every length, 1e-2 A^2 on every squared length, 1e-3 on the L1 cutoff distance (as in the
// fp64 screen's box stage), 1e-3 relative on the half-plane and angle pretests.
#pragma once

#include "classify_screen.cuh"

namespace fr3d {

constexpr int S32_REC = 80;
#ifdef FR3D_S32_SCALAR_ROWS
constexpr int S32_STRIDE = 49;                  // odd: "lane k reads its own row" is bank-conflict free
#else
// rows are copied in 16-byte pieces and read with 128-bit loads: 13 x 16 B per row (odd: a quarter warp's 128-bit
// reads of "its own rows" touch 8 different 16-byte bank groups)
constexpr int S32_STRIDE = 52;
#endif
constexpr int S32_BB = 9, S32_MASKS = 30, S32_BASE = 32;
constexpr double S32_FX = 65536.0;
constexpr double S32_RANGE = 32000.0;
constexpr float S32_FAR = 512.f;
#ifndef FR3D_S32_BULK
// 1: every lane stages its row with ONE bulk copy (cp.async.bulk.shared.global + the warp's mbarrier, 128 B then 192 B)
// instead of 8 + 12 16-byte cp.async and as many shuffles.  Correct (same hits) but measured SLOWER on the bench
// workload - classify 0.775 vs 0.759 ms (profiles/ab_bulk_copy_r2.log): the copy engine's per-request cost outweighs
// the ~35 issue slots saved per pair when a request is only 128-192 bytes.  Kept as a build option, default off.
#define FR3D_S32_BULK 0
#endif
#ifndef FR3D_S32_WARPS
#define FR3D_S32_WARPS 24      // 24 (80 registers, no spill) and 32 (64 registers) warps per SM measure the same
#endif
constexpr int S32_WARPS = FR3D_S32_WARPS;
constexpr int S32_FIXED = (SCR_UBOX * 8 + SCR_MAX_BOXES * SCR_BOX_STRIDE + SCR_HULL + SCR_UBOX * 2 + 2 * (FR3D_N_PARENTS + 1) + 3) / 4 * 4;
constexpr size_t S32_SMEM = sizeof(float) * (S32_FIXED + S32_WARPS * 32 * S32_STRIDE);

// Builds the screen records: one block per 32 consecutive nts.  The four source planes of those nts are contiguous
// in memory, so they are staged in shared memory with coalesced asynchronous copies (every load independent of
// every other), then the 32 x 80 output words are written coalesced, each computed from the staged doubles.
constexpr int SREC_NTS = 32, SREC_THREADS = 320;          // 320 = 4 x S32_REC
constexpr int SREC_C = 0, SREC_U = SREC_C + SREC_NTS * 3, SREC_BB = SREC_U + SREC_NTS * 9,
              SREC_BASE = SREC_BB + SREC_NTS * FR3D_BB_SLOTS * 3, SREC_DOUBLES = SREC_BASE + SREC_NTS * FR3D_BASE_SLOTS * 3;

__global__ void __launch_bounds__(SREC_THREADS) screen_records_kernel(fr3d_nts_t nts, float *__restrict__ rec, int64_t *counters) {
    __shared__ __align__(16) double sd[SREC_DOUBLES];
    __shared__ uint32_t sm[2 * SREC_NTS];
    const int n0 = blockIdx.x * SREC_NTS;
    const int cnt = min(SREC_NTS, nts.n_nt - n0);
    if (cnt <= 0) return;
    const int tid = threadIdx.x;
    const bool aligned16 = (((size_t)nts.center | (size_t)nts.rot | (size_t)nts.bb_xyz | (size_t)nts.base_xyz) & 15) == 0;
    if (cnt == SREC_NTS && aligned16) {   // every plane segment of a full block then starts on a 16-byte boundary
        for (int t = tid; t < SREC_NTS * 3 / 2; t += SREC_THREADS) cp_async16(sd + SREC_C + 2 * t, nts.center + (size_t)n0 * 3 + 2 * t);
        for (int t = tid; t < SREC_NTS * 9 / 2; t += SREC_THREADS) cp_async16(sd + SREC_U + 2 * t, nts.rot + (size_t)n0 * 9 + 2 * t);
        for (int t = tid; t < SREC_NTS * FR3D_BB_SLOTS * 3 / 2; t += SREC_THREADS)
            cp_async16(sd + SREC_BB + 2 * t, nts.bb_xyz + (size_t)n0 * FR3D_BB_SLOTS * 3 + 2 * t);
        for (int t = tid; t < SREC_NTS * FR3D_BASE_SLOTS * 3 / 2; t += SREC_THREADS)
            cp_async16(sd + SREC_BASE + 2 * t, nts.base_xyz + (size_t)n0 * FR3D_BASE_SLOTS * 3 + 2 * t);
    } else {
        for (int t = tid; t < cnt * 3; t += SREC_THREADS) cp_async8(sd + SREC_C + t, nts.center + (size_t)n0 * 3 + t);
        for (int t = tid; t < cnt * 9; t += SREC_THREADS) cp_async8(sd + SREC_U + t, nts.rot + (size_t)n0 * 9 + t);
        for (int t = tid; t < cnt * FR3D_BB_SLOTS * 3; t += SREC_THREADS)
            cp_async8(sd + SREC_BB + t, nts.bb_xyz + (size_t)n0 * FR3D_BB_SLOTS * 3 + t);
        for (int t = tid; t < cnt * FR3D_BASE_SLOTS * 3; t += SREC_THREADS)
            cp_async8(sd + SREC_BASE + t, nts.base_xyz + (size_t)n0 * FR3D_BASE_SLOTS * 3 + t);
    }
    if (tid < cnt) cp_async4(sm + tid, nts.base_mask + n0 + tid);
    else if (tid >= SREC_NTS && tid < SREC_NTS + cnt) cp_async4(sm + tid, nts.bb_mask + n0 + (tid - SREC_NTS));
    cp_async_wait_all();
    __syncthreads();
    // each thread owns ONE word position w of the record (320 threads = 4 x 80) and walks the nts k = tid / 80, + 4, ...:
    // what the word is made of is decided once, the loop body is two shared-memory loads, a subtract and a store.
    // The five special words (fixed-point centre, masks) are written after the loop.
    const int w = tid % S32_REC;
    int src = -1, stride = 0, csub = -1, slot = -1;    // value = sd[src + k * stride] - (csub >= 0 ? sd[csub + 3 k] : 0)
    if (w >= 3 && w < 9) { src = SREC_U + ((w - 3) % 3) * 3 + (w < 6 ? 0 : 2); stride = 9; }
    else if (w >= S32_BB && w < S32_MASKS) {
        const int b = (w - S32_BB) / 3, comp = (w - S32_BB) - b * 3;
        src = SREC_BB + (b == 6 ? 9 : b + 1) * 3 + comp; stride = FR3D_BB_SLOTS * 3; csub = SREC_C + comp;
    } else if (w >= S32_BASE) {
        src = SREC_BASE + (w - S32_BASE); stride = FR3D_BASE_SLOTS * 3; csub = SREC_C + (w - S32_BASE) % 3;
        slot = (w - S32_BASE) / 3;
    }
    float *out = rec + (size_t)n0 * S32_REC;
    // The error bounds of the single-precision kernels assume atoms within S32_FAR of their base centre (real bases
    // stay within 5 A).  A record with a present atom beyond that gets FR3D_FLAG_DUP's bit in ITS copy of base_mask,
    // which sends its stacking pairs to the fp64 kernel (classify_stack32.cuh); the screens need no such guard: an
    // atom that far cannot pass a proximity test by any margin.
    __shared__ uint32_t s_far[SREC_NTS];
    if (tid < SREC_NTS) s_far[tid] = 0;
    __syncthreads();
    if (src >= 0) {
#pragma unroll 4
        for (int k = tid / S32_REC; k < cnt; k += SREC_THREADS / S32_REC) {
            const uint32_t bm = sm[k];
            const double x = sd[src + k * stride];
            float v = (float)(csub >= 0 ? x - sd[csub + k * 3] : x);
            if (!(bm & FR3D_MASK_HAS_FRAME)) v = 0.f;          // never on a candidate pair; only the masks are kept
            else if (slot >= 0 && ((bm >> slot) & 1u) && !(fabsf(v) < S32_FAR)) atomicOr(&s_far[k], 1u << 22);
            out[k * S32_REC + w] = v;
        }
    }
    __syncthreads();
    if (tid < cnt) out[tid * S32_REC + S32_MASKS] = __uint_as_float(sm[tid] | s_far[tid]);
    else if (tid >= SREC_NTS && tid < SREC_NTS + cnt) out[(tid - SREC_NTS) * S32_REC + S32_MASKS + 1] = __uint_as_float(sm[tid]);
    else if (tid >= 2 * SREC_NTS && tid < 2 * SREC_NTS + 3 * cnt) {        // the centre as 2^-16 A fixed point
        const int t = tid - 2 * SREC_NTS, k = t / 3, comp = t - k * 3;
        float v = 0.f;
        if (sm[k] & FR3D_MASK_HAS_FRAME) {
            const double x = sd[SREC_C + t];
            if (!(fabs(x) < S32_RANGE)) atomicOr((unsigned long long *)&counters[2], 1ull);      // FR3D_E_RANGE
            v = __int_as_float((int)rint(fmax(-S32_RANGE, fmin(S32_RANGE, x)) * S32_FX));
        }
        out[k * S32_REC + comp] = v;
    }
}

struct SFrame {
    float a0, a1, a2;        // x axis of the base frame  (first column of rotation_matrix)
    float n0, n1, n2;        // base normal               (third column)
};

// which of the 16 points q (relative to their own centre; + (e0, e1, e2) = relative to the centre of frame f)
// lie in the cylinder |z| < 4.5, x^2 + y^2 < r2max of frame f; converged, branch free
__device__ __forceinline__ unsigned s32_cylinder_mask(const SFrame &f, const float *q, float e0, float e1, float e2,
                                                      float r2max) {
    unsigned cand = 0;
#pragma unroll
    for (int a = 0; a < FR3D_BASE_SLOTS; ++a) {
        const float v0 = q[a * 3] + e0, v1 = q[a * 3 + 1] + e1, v2 = q[a * 3 + 2] + e2;
        const float z = v0 * f.n0 + v1 * f.n1 + v2 * f.n2;
        const float r2 = (v0 * v0 + v1 * v1 + v2 * v2) - z * z;
        if (fabsf(z) < 4.501f && r2 < r2max) cand |= 1u << a;
    }
    return cand;
}

// any of the candidate points inside the hull prism of f's base (half planes carry their margin already)?
__device__ __forceinline__ bool s32_hull_any(const SFrame &f, const float *hull, const float *q, float e0, float e1,
                                             float e2, unsigned cand) {
    const float b0 = f.n1 * f.a2 - f.n2 * f.a1, b1 = f.n2 * f.a0 - f.n0 * f.a2, b2 = f.n0 * f.a1 - f.n1 * f.a0;   // y axis
    bool any = false;
    while (cand && !any) {
        const int a = __ffs(cand) - 1;
        cand &= cand - 1;
        const float v0 = q[a * 3] + e0, v1 = q[a * 3 + 1] + e1, v2 = q[a * 3 + 2] + e2;
        const float x = v0 * f.a0 + v1 * f.a1 + v2 * f.a2, y = v0 * b0 + v1 * b1 + v2 * b2;
        bool in = true;
#pragma unroll
        for (int k = 0; k < 7; ++k) in = in && (hull[k * 3] * x + hull[k * 3 + 1] * y + hull[k * 3 + 2] > 0.f);
        any = in;
    }
    return any;
}

// any backbone oxygen (record words 9..26: slots 1..6) inside the slab / disc that contains every ring and
// near-ring ellipse of f's base?
__device__ __forceinline__ bool s32_so_possible(const SFrame &f, const float *bb, uint32_t bbmask, float e0, float e1,
                                                float e2, float r2max) {
    bool any = false;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const float v0 = bb[k * 3] + e0, v1 = bb[k * 3 + 1] + e1, v2 = bb[k * 3 + 2] + e2;
        const float z = v0 * f.n0 + v1 * f.n1 + v2 * f.n2;
        const float r2 = (v0 * v0 + v1 * v1 + v2 * v2) - z * z;
        any = any || (((bbmask >> (k + 1)) & 1u) && fabsf(z) < 3.601f && r2 < r2max);
    }
    return any;
}

__global__ void __launch_bounds__(S32_WARPS * 32, 1)
classify_screen32_kernel(const float *__restrict__ rec, const fr3d_box_t *__restrict__ boxes,
                         const int32_t *__restrict__ box_index, const int32_t *__restrict__ pair_i,
                         const int32_t *__restrict__ pair_j, int64_t pair_cap, uint32_t cats,
                         int32_t *__restrict__ wl_light, int32_t *__restrict__ wl_stack, int32_t *__restrict__ wl_pair,
                         int32_t *__restrict__ wl_bph, uint32_t *__restrict__ wl_bph_combos, int64_t *counters) {
    extern __shared__ __align__(16) float s32_smem[];
    float *ubox = s32_smem;                                          // [SCR_UBOX][8] union boxes
    float *sbox = ubox + SCR_UBOX * 8;                               // [SCR_MAX_BOXES][SCR_BOX_STRIDE] mid, half
    float *shull = sbox + SCR_MAX_BOXES * SCR_BOX_STRIDE;            // [parents][7][3], margin folded into [2]
    int *sidx = reinterpret_cast<int *>(shull + SCR_HULL);           // [SCR_UBOX][2] start, count
    float *s_r2 = reinterpret_cast<float *>(sidx + SCR_UBOX * 2);    // [parent + 1]: sO disc, then hull disc
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *rows = s32_smem + S32_FIXED + warp * 32 * S32_STRIDE;
    __shared__ int s_nboxes;
#if FR3D_S32_BULK && !defined(FR3D_S32_SCALAR_ROWS)
    __shared__ __align__(8) unsigned long long s_bar[S32_WARPS];       // one mbarrier per warp: its rows are private
    if (lane == 0) mbar_init(&s_bar[warp], 1);
    fence_proxy_async_smem();
    uint32_t bar_parity = 0;
#endif
    if (threadIdx.x == 0) s_nboxes = 0;
    if (threadIdx.x <= FR3D_N_PARENTS) {
        const int p = (int)threadIdx.x - 1;
        s_r2[threadIdx.x] = p < 0 ? 0.f : (float)c_tab.so_r2max[p] + 1e-2f;
        s_r2[FR3D_N_PARENTS + 1 + threadIdx.x] = p < 0 ? 0.f : (float)c_tab.hull_r2max[p] + 1e-2f;
    }
    __syncthreads();
    // prologue: union of the cutoff boxes of every (combination, sign) entry (see classify_screen.cuh)
    for (int item = threadIdx.x; item < SCR_UBOX * 8; item += blockDim.x) {
        const int e = item >> 3, q = item & 7;
        const int bstart = box_index[e * 2], bcount = box_index[e * 2 + 1];
        const bool is_min = !(q & 1);
        double acc = is_min ? 1e30 : -1e30;
        for (int b = 0; b < bcount; ++b) {
            const double v = reinterpret_cast<const double *>(boxes + bstart + b)[q];
            acc = is_min ? fmin(acc, v) : fmax(acc, v);
        }
        ubox[item] = (float)acc;
        if (q == 0) { sidx[e * 2] = bstart; sidx[e * 2 + 1] = bcount; if (bcount > 0) atomicMax(&s_nboxes, bstart + bcount); }
    }
    for (int item = threadIdx.x; item < SCR_HULL; item += blockDim.x) {
        const double *l = &c_tab.hull[0][0][0] + (item / 3) * 3;
        const double v = (&c_tab.hull[0][0][0])[item];
        shull[item] = (float)(item % 3 == 2 ? v + 1e-3 * (fabs(l[0]) + fabs(l[1])) + 1e-5 * fabs(v) + 1e-6 : v);
    }
    __syncthreads();
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
            if ((item & 3) == 0) {
                sbox[(item >> 2) * SCR_BOX_STRIDE + 8] = (float)bx[8];
                sbox[(item >> 2) * SCR_BOX_STRIDE + 9] = (float)bx[9];
            }
        }
    __syncthreads();
    long long n_pairs = counters[0];
    if (n_pairs > pair_cap) n_pairs = pair_cap;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long rounds = (n_pairs + stride - 1) / stride;
    int nx1 = first < n_pairs ? pair_i[first] : 0, nx2 = first < n_pairs ? pair_j[first] : 0;
    constexpr float INV_FX = (float)(1.0 / S32_FX);
#pragma unroll 1
    for (long long r = 0; r < rounds; ++r) {
        const long long p = first + r * stride;
        const bool valid = p < n_pairs;
        const int n1 = nx1, n2 = nx2;
        nx1 = p + stride < n_pairs ? pair_i[p + stride] : 0;
        nx2 = p + stride < n_pairs ? pair_j[p + stride] : 0;
        if (!__any_sync(FULL, valid)) continue;
        // phase A: words 0..31 of the 32 nt2 records, one coalesced 128-byte copy per row (row k belongs to lane k)
        __syncwarp();
#ifdef FR3D_S32_SCALAR_ROWS
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            const int nb = __shfl_sync(FULL, n2, k);
            cp_async4(rows + k * S32_STRIDE + lane, rec + (size_t)nb * S32_REC + lane);
        }
#elif FR3D_S32_BULK
        // every lane fetches ITS row with one bulk copy (128 B); the warp's mbarrier counts the 32 x 128 bytes
        fence_proxy_async_smem();
        if (lane == 0) mbar_arrive_expect_tx(&s_bar[warp], 32u * 128u);
        bulk_copy_g2s(rows + lane * S32_STRIDE, rec + (size_t)n2 * S32_REC, 128u, &s_bar[warp]);
#else
#pragma unroll
        for (int j = 0; j < 8; ++j) {                       // 8 lanes x 16 B per row, four rows per instruction
            const int row = j * 4 + (lane >> 3), piece = (lane & 7) * 4;
            const int nb = __shfl_sync(FULL, n2, row);
            cp_async16(rows + row * S32_STRIDE + piece, rec + (size_t)nb * S32_REC + piece);
        }
#endif
        // nt1: read directly (pairs come grouped by nt1, so these loads are near-broadcast)
        const float *A = rec + (size_t)n1 * S32_REC;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A + 32));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A + 64));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A + 79));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + (size_t)nx1 * S32_REC));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + (size_t)nx1 * S32_REC + 31));
        const int c1x = __float_as_int(A[0]), c1y = __float_as_int(A[1]), c1z = __float_as_int(A[2]);
        SFrame f1;
        f1.a0 = A[3]; f1.a1 = A[4]; f1.a2 = A[5]; f1.n0 = A[6]; f1.n1 = A[7]; f1.n2 = A[8];
        const uint32_t bm1 = __float_as_uint(A[S32_MASKS]), bb1 = __float_as_uint(A[S32_MASKS + 1]);
        const int par1 = (int)((bm1 >> 16) & 15u) - 1;
        const bool par1_ok = par1 >= 0 && par1 < FR3D_N_PARENTS;
        unsigned fl = 0, live = 0, bph_combos = 0;
        float dX0 = 0, dY0 = 0, dZ0 = 0, dX1 = 0, dY1 = 0, dZ1 = 0, nZ = 0, ang0 = 0.f, ang1 = 0.f;
        int entry0 = 0, entry1 = 0;
#if FR3D_S32_BULK && !defined(FR3D_S32_SCALAR_ROWS)
        mbar_wait_parity(&s_bar[warp], bar_parity);
        bar_parity ^= 1u;
#else
        cp_async_wait_all();
        __syncwarp();
#endif
#ifdef FR3D_S32_SCALAR_ROWS
        const float *R = rows + lane * S32_STRIDE;
#else
        float R[S32_BASE];                                  // this lane's row in registers: eight 128-bit loads
        {
            const float4 *R4 = reinterpret_cast<const float4 *>(rows + lane * S32_STRIDE);
#pragma unroll
            for (int j = 0; j < S32_BASE / 4; ++j) {
                const float4 t = R4[j];
                R[4 * j] = t.x; R[4 * j + 1] = t.y; R[4 * j + 2] = t.z; R[4 * j + 3] = t.w;
            }
        }
#endif
        SFrame f2;
        f2.a0 = f2.a1 = f2.a2 = f2.n0 = f2.n1 = f2.n2 = 0.f;
        float e0 = 0.f, e1 = 0.f, e2 = 0.f;
        uint32_t bm2 = 0, bb2 = 0;
        int par2 = -1;
        bool par2_ok = false;
        if (valid) {
            // e = centre of nt2 - centre of nt1
            e0 = (float)(__float_as_int(R[0]) - c1x) * INV_FX;
            e1 = (float)(__float_as_int(R[1]) - c1y) * INV_FX;
            e2 = (float)(__float_as_int(R[2]) - c1z) * INV_FX;
            f2.a0 = R[3]; f2.a1 = R[4]; f2.a2 = R[5]; f2.n0 = R[6]; f2.n1 = R[7]; f2.n2 = R[8];
            bm2 = __float_as_uint(R[S32_MASKS]); bb2 = __float_as_uint(R[S32_MASKS + 1]);
            par2 = (int)((bm2 >> 16) & 15u) - 1;
            par2_ok = par2 >= 0 && par2 < FR3D_N_PARENTS;
            if (cats & FR3D_CAT_SO) {
                if (par1_ok && s32_so_possible(f1, R + S32_BB, bb2, e0, e1, e2, s_r2[par1 + 1])) fl |= SCR_SO12;
                if (par2_ok && s32_so_possible(f2, A + S32_BB, bb1, -e0, -e1, -e2, s_r2[par2 + 1])) fl |= SCR_SO21;
            }
            if ((cats & FR3D_CAT_SUGAR_RIBOSE) && par1 >= 0 && par2 >= 0 && par1 != 4 && par2 != 4 &&
                (bb1 >> 6 & 1u) && (bb2 >> 6 & 1u)) {
                const float o0 = A[S32_BB + 15] - (R[S32_BB + 15] + e0), o1 = A[S32_BB + 16] - (R[S32_BB + 16] + e1),
                            o2 = A[S32_BB + 17] - (R[S32_BB + 17] + e2);
                if (o0 * o0 + o1 * o1 + o2 * o2 < 14.4401f + 1e-2f) fl |= SCR_SR;
            }
            if ((cats & FR3D_CAT_BACKBONE) && par1_ok) {
                // some massive atom of a base-phosphate site of nt1 within 4.5 A of a backbone oxygen of nt2
                // (bit site * 6 + i of bph_combos, i in the order of the fp64 rule: O5' OP1 OP2 prevO3' O2' O4')
                const int nsites = c_tab.n_sites[par1];
#pragma unroll 1
                for (int s = 0; s < nsites; ++s) {
                    const int ms = c_tab.site_m[par1][s], hs = c_tab.site_h[par1][s];
                    if (!((bm1 >> ms) & 1u) || !((bm1 >> hs) & 1u)) continue;        // NA:1937: both atoms needed
                    // in nt2's centre-relative coordinates
                    const float m0 = A[S32_BASE + ms * 3] - e0, m1 = A[S32_BASE + ms * 3 + 1] - e1,
                                m2 = A[S32_BASE + ms * 3 + 2] - e2;
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        const int k = i < 4 ? i : i + 1;                 // record slots OP1 OP2 O5' O4' | O2' prevO3'
                        const int oslot = k == 6 ? 9 : k + 1;
                        const float o0 = R[S32_BB + k * 3], o1 = R[S32_BB + k * 3 + 1], o2 = R[S32_BB + k * 3 + 2];
                        const float q0 = m0 - o0, q1 = m1 - o1, q2 = m2 - o2;
                        if (((bb2 >> oslot) & 1u) && (q0 * q0 + q1 * q1 + q2 * q2 < 20.2501f + 1e-2f)) {
                            // rare: every label also needs the angle massive-H-oxygen above 110 degrees; keep the
                            // pair unless it is certainly below 107 (cos >= 0 or sin^2 > 10.3 cos^2, tan^2 72.7 = 10.3)
                            const float h0 = A[S32_BASE + hs * 3] - e0, h1 = A[S32_BASE + hs * 3 + 1] - e1,
                                        h2 = A[S32_BASE + hs * 3 + 2] - e2;
                            const float u0 = m0 - h0, u1 = m1 - h1, u2 = m2 - h2;
                            const float w0 = o0 - h0, w1 = o1 - h1, w2 = o2 - h2;
                            const float cs = u0 * w0 + u1 * w1 + u2 * w2;
                            const float x0 = u1 * w2 - u2 * w1, x1 = u2 * w0 - u0 * w2, x2 = u0 * w1 - u1 * w0;
                            if (!(cs >= 1e-3f || x0 * x0 + x1 * x1 + x2 * x2 > 10.3f * 1.001f * cs * cs + 1e-3f))
                                bph_combos |= 1u << (s * 6 + ((0x345021 >> (4 * i)) & 15));
                        }
                    }
                }
                if (bph_combos) fl |= SCR_BPH;
            }
        }
        // phase B: the base points of the 32 nt2 records take the place of the phase A rows
        __syncwarp();
#ifdef FR3D_S32_SCALAR_ROWS
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            const int nb = __shfl_sync(FULL, n2, k);
            const float *src = rec + (size_t)nb * S32_REC + S32_BASE + lane;
            float *dst = rows + k * S32_STRIDE + lane;
            cp_async4(dst, src);
            if (lane < 16) cp_async4(dst + 32, src + 32);
        }
#elif FR3D_S32_BULK
        fence_proxy_async_smem();                           // the lanes have read their phase A rows (__syncwarp above)
        if (lane == 0) mbar_arrive_expect_tx(&s_bar[warp], 32u * 192u);
        bulk_copy_g2s(rows + lane * S32_STRIDE, rec + (size_t)n2 * S32_REC + S32_BASE, 192u, &s_bar[warp]);
        mbar_wait_parity(&s_bar[warp], bar_parity);
        bar_parity ^= 1u;
#else
#pragma unroll
        for (int j = 0; j < 12; ++j) {                      // 12 pieces per row: 4 lanes x 16 B on eight rows per instruction
            const int row = (j & 3) * 8 + (lane >> 2), piece = ((j >> 2) * 4 + (lane & 3)) * 4;
            const int nb = __shfl_sync(FULL, n2, row);
            cp_async16(rows + row * S32_STRIDE + piece, rec + (size_t)nb * S32_REC + S32_BASE + piece);
        }
#endif
#if !(FR3D_S32_BULK && !defined(FR3D_S32_SCALAR_ROWS))
        cp_async_wait_all();
        __syncwarp();
#endif
        if (valid) {
#ifdef FR3D_S32_SCALAR_ROWS
            const float *RB = rows + lane * S32_STRIDE;
#else
            float RB[FR3D_BASE_SLOTS * 3];                  // nt2's 16 base points in registers: twelve 128-bit loads
            {
                const float4 *R4 = reinterpret_cast<const float4 *>(rows + lane * S32_STRIDE);
#pragma unroll
                for (int j = 0; j < FR3D_BASE_SLOTS * 3 / 4; ++j) {
                    const float4 t = R4[j];
                    RB[4 * j] = t.x; RB[4 * j + 1] = t.y; RB[4 * j + 2] = t.z; RB[4 * j + 3] = t.w;
                }
            }
#endif
            const float *Ab = A + S32_BASE;
            if ((cats & FR3D_CAT_STACKING) && par1_ok && par2_ok) {
                const unsigned c12 = s32_cylinder_mask(f1, RB, e0, e1, e2, s_r2[FR3D_N_PARENTS + 2 + par1]) & bm2;
                const unsigned c21 = s32_cylinder_mask(f2, Ab, -e0, -e1, -e2, s_r2[FR3D_N_PARENTS + 2 + par2]) & bm1;
                // (the hull test indexes the points with a run-time slot: it reads the shared-memory row)
                if (s32_hull_any(f1, shull + par1 * 21, rows + lane * S32_STRIDE, e0, e1, e2, c12) ||
                    s32_hull_any(f2, shull + par2 * 21, Ab, -e0, -e1, -e2, c21)) fl |= SCR_STACK;
            }
            if ((bm1 & 1u) && (bm2 & 1u) && par1_ok && par2_ok) {
                unsigned ori_ok = 0;
                if (c_tab.combo_ok[par1 * FR3D_N_PARENTS + par2]) ori_ok |= 1u;
                if (c_tab.combo_ok[par2 * FR3D_N_PARENTS + par1]) ori_ok |= 2u;
                nZ = f1.n0 * f2.n0 + f1.n1 * f2.n1 + f1.n2 * f2.n2;
                if ((cats & FR3D_CAT_COPLANAR) && ori_ok && fabsf(nZ) > 0.7757f - 1e-4f) {
                    const float lim = (0.3381f * 0.3381f) * 1.001f * (e0 * e0 + e1 * e1 + e2 * e2) + 1e-4f;
                    const float d1 = e0 * f1.n0 + e1 * f1.n1 + e2 * f1.n2;
                    const float d2 = e0 * f2.n0 + e1 * f2.n1 + e2 * f2.n2;
                    if (d1 * d1 < lim && d2 * d2 < lim) fl |= SCR_CP;
                }
                // glycosidic atom of nt2 minus glycosidic atom of nt1
                const float gd0 = (RB[0] + e0) - Ab[0], gd1 = (RB[1] + e1) - Ab[1], gd2 = (RB[2] + e2) - Ab[2];
                const int sg = nZ > 0 ? 0 : 1;
                entry0 = (par1 * FR3D_N_PARENTS + par2) * 2 + sg;
                entry1 = (par2 * FR3D_N_PARENTS + par1) * 2 + sg;
                // y axes of the two frames
                const float b10 = f1.n1 * f1.a2 - f1.n2 * f1.a1, b11 = f1.n2 * f1.a0 - f1.n0 * f1.a2, b12 = f1.n0 * f1.a1 - f1.n1 * f1.a0;
                const float b20 = f2.n1 * f2.a2 - f2.n2 * f2.a1, b21 = f2.n2 * f2.a0 - f2.n0 * f2.a2, b22 = f2.n0 * f2.a1 - f2.n1 * f2.a0;
                const float m11 = b10 * b20 + b11 * b21 + b12 * b22;
                // the sign of nZ picks the half of the cutoff table (NA:2541): where single precision cannot be trusted
                // with it, the pair simply goes on to the fp64 list kernel
                if (ori_ok && fabsf(nZ) < 1e-4f) { fl |= SCR_BP; ori_ok = 0; }
#pragma unroll
                for (int ori = 0; ori < 2; ++ori) {
                    const SFrame &fa = ori ? f2 : f1;
                    const float s = ori ? -1.f : 1.f;
                    const float v0 = s * gd0, v1 = s * gd1, v2 = s * gd2;
                    const float ya0 = ori ? b20 : b10, ya1 = ori ? b21 : b11, ya2 = ori ? b22 : b12;
                    const float yb0 = ori ? b10 : b20, yb1 = ori ? b11 : b21, yb2 = ori ? b12 : b22;
                    const float dZ = v0 * fa.n0 + v1 * fa.n1 + v2 * fa.n2;
                    const float dX = v0 * fa.a0 + v1 * fa.a1 + v2 * fa.a2;
                    const float dY = v0 * ya0 + v1 * ya1 + v2 * ya2;
                    if (ori) { dX1 = dX; dY1 = dY; dZ1 = dZ; } else { dX0 = dX; dY0 = dY; dZ0 = dZ; }
                    const float *ub = ubox + (ori ? entry1 : entry0) * 8;
                    const float cd = fmaxf(0.f, ub[0] - dX) + fmaxf(0.f, dX - ub[1]) + fmaxf(0.f, ub[2] - dY) +
                                     fmaxf(0.f, dY - ub[3]) + fmaxf(0.f, ub[4] - dZ) + fmaxf(0.f, dZ - ub[5]) +
                                     3.f * (fmaxf(0.f, ub[6] - nZ) + fmaxf(0.f, nZ - ub[7]));
                    if (((ori_ok >> ori) & 1u) && !(fabsf(dZ) > 3.601f) && cd < 1.0f + 1e-3f) {             // NA:2380
                        live |= 1u << ori;
                        // in-plane angle of this orientation (NA:2594): atan2(M[1][1], M[0][1]), M = Ua^T Ub
                        const float m01 = fa.a0 * yb0 + fa.a1 * yb1 + fa.a2 * yb2;
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
            const int src_a = __ffs(pending) - 1;
            const unsigned rest = pending & (pending - 1);
            const int src_b = (two_tasks && rest) ? __ffs(rest) - 1 : -1;
            const bool second = two_tasks && (lane >> 4);
            const int srcl = second ? src_b : src_a;
            const int sub = two_tasks ? (lane & 15) : lane;
            const bool first_ori = live & 1u;
            const int from = srcl < 0 ? 0 : srcl;
            const float tX = __shfl_sync(FULL, first_ori ? dX0 : dX1, from);
            const float tY = __shfl_sync(FULL, first_ori ? dY0 : dY1, from);
            const float tZ = __shfl_sync(FULL, first_ori ? dZ0 : dZ1, from);
            const float tN = __shfl_sync(FULL, nZ, from);
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
                // the angle term (NA:2596-2611; ranges with amin > amax wrap around).  The angle lives in (-90, 270]:
                // within 0.1 degree of that seam single precision may land on the other side, so both images count
                const float tB = tA + (tA < -89.9f ? 360.f : (tA > 269.9f ? -360.f : 0.f));
                const float below = fmaxf(0.f, amin - tA), above = fmaxf(0.f, tA - amax);
                const float below2 = fmaxf(0.f, amin - tB), above2 = fmaxf(0.f, tB - amax);
                cd += 0.05f * (amin < amax ? fminf(below + above, below2 + above2) : fminf(fminf(below, above), fminf(below2, above2)));
                hit = cd < 1.0f + 2e-3f;
            }
            const unsigned hits = __ballot_sync(FULL, hit);
            if (lane == src_a || lane == src_b) {
                const bool any = two_tasks ? ((lane == src_b ? hits >> 16 : hits & 0xFFFFu) != 0) : hits != 0;
                if (any) { fl |= SCR_BP; live = 0; }
                else live &= live - 1;
            }
            pending = __ballot_sync(FULL, live != 0);
        }
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
        if (want_bph) {
            wl_bph[base_b + __popc(bb & below)] = (int32_t)p;
            wl_bph_combos[base_b + __popc(bb & below)] = bph_combos;
        }
    }
}

}  // namespace fr3d
