// K4, stacking work list in single precision with an exact double-precision fallback.
//
// check_base_base_stacking (NA:1602-1739) produces a LABEL only (s35 / s53 / s33 / s55, near or true) - no number of
// it reaches the output.  Every step of the rule is a comparison of a geometric quantity with a threshold (atom
// heights against 0 and 4.5, hull half planes against 0, the normals' product against 0.5 and 0.6, the minimum
// atom-atom distance against 1 and 4).  So the rule can be evaluated in single precision on the screen records
// (classify_screen32.cuh) as long as every comparison is CERTAIN: the quantity is further from its threshold than a
// margin far above the fp32 error (heights and distances: 2e-3 A against an error below 1e-4 A; half planes and
// normals: see below).  A pair with any uncertain comparison - and any pair with a "twin" slot (modified residues
// with ideal-position copies) - is appended to a fallback list that an fp64 kernel then processes, so the labels are
// exactly those of the fp64 path; about 0.5 % of the listed pairs go that way.
// ncu on the fp64 kernel (profiles/step_r1n_*): 168 registers, 11 warps per SM, 38 % issue utilisation, fp64
// dependency latency, 226 warp instructions per listed pair; here: fp32 latencies, 110 instructions, 12 warps per SM.
//
// The minimum distance (15 x 15 atom pairs) is needed only for "mind < 1" (no label) and, for a true stack,
// "mind < 4".  "mind > 1" is certified without the pair loop from heights: every atom of nt2 is at least h2 = min |z|
// above base 1's plane and every atom of nt1 at most t1 = max |z| off its own plane, so mind >= h2 - t1 (and the same
// with the roles swapped); the pair loop runs only when that bound fails or when a true stack needs "mind < 4", and
// in the latter case it stops at the first atom pair closer than 4 - margin.
#pragma once

#include "classify_screen32.cuh"

namespace fr3d {

#ifndef FR3D_ST32_WARPS
#define FR3D_ST32_WARPS 12      // 12 / 16 / 20 warps: classify 0.843 / 0.850 / 0.873 ms (fewer warps leave more L1 for nt1's records)
#endif
constexpr int ST32_WARPS = FR3D_ST32_WARPS;
constexpr int ST32_STRIDE = 84;                  // 21 x 16 B (odd): rows are copied in 16-byte pieces and read with 128-bit loads
constexpr int ST32_FIXED = (SCR_HULL + FR3D_N_PARENTS * 7 + FR3D_N_PARENTS + 3) / 4 * 4;        // hull planes, their margins, r2in
constexpr size_t ST32_SMEM = sizeof(float) * (ST32_FIXED + ST32_WARPS * 32 * ST32_STRIDE);
constexpr float ST32_DZ = 2e-3f;                 // margin on heights and distances (A)

// return_overlap (NA:1530-1571) of the 16 points q (+ e = relative to f's centre) over f's base, with certainty:
// 0 certainly no overlap, 1 certainly overlap, 2 cannot tell in single precision.  zsel = height of the atom
// closest to the plane (asel = its slot); when the result is 1 every atom is on zsel's side by more than the margin.
__device__ __forceinline__ int st32_overlap(const SFrame &f, float b0, float b1, float b2, const float *hull,
                                            const float *hmargin, float r2in_sure, const float *q, const float *qdyn,
                                            float e0, float e1, float e2, uint32_t bm, float &zsel, int &asel) {
    // q: the points for the unrolled pass (may live in registers); qdyn: the same points in memory, for the hull loop
    // that indexes them with a run-time slot
    unsigned cand = 0;
    bool in_sure = false;
    float best = 1e30f, zmin = 1e30f, zmax = -1e30f;
    zsel = 0.f;
    asel = 0;
#pragma unroll
    for (int a = 0; a < FR3D_BASE_SLOTS; ++a) {
        const bool have = (bm >> a) & 1u;
        const float v0 = q[a * 3] + e0, v1 = q[a * 3 + 1] + e1, v2 = q[a * 3 + 2] + e2;
        const float z = v0 * f.n0 + v1 * f.n1 + v2 * f.n2;
        const float r2 = (v0 * v0 + v1 * v1 + v2 * v2) - z * z;
        const float az = fabsf(z);
        if (have && az < 4.5f + ST32_DZ && r2 < 16.01f) cand |= 1u << a;
        if (have && az < 4.5f - ST32_DZ && r2 < r2in_sure) in_sure = true;
        if (have && az < best) { best = az; zsel = z; asel = a; }
        if (have) { zmin = fminf(zmin, z); zmax = fmaxf(zmax, z); }
    }
    // atoms on both sides of the plane (NA:1567)?
    const bool pos_sure = zmax > ST32_DZ, neg_sure = zmin < -ST32_DZ, pos_maybe = zmax > -ST32_DZ, neg_maybe = zmin < ST32_DZ;
    bool in_maybe = in_sure;
    while (cand && !in_sure) {
        const int a = __ffs(cand) - 1;
        cand &= cand - 1;
        const float v0 = qdyn[a * 3] + e0, v1 = qdyn[a * 3 + 1] + e1, v2 = qdyn[a * 3 + 2] + e2;
        const float x = v0 * f.a0 + v1 * f.a1 + v2 * f.a2, y = v0 * b0 + v1 * b1 + v2 * b2;
        const float az = fabsf(v0 * f.n0 + v1 * f.n1 + v2 * f.n2);
        bool sure = az < 4.5f - ST32_DZ, maybe = true;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const float h = hull[k * 3] * x + hull[k * 3 + 1] * y + hull[k * 3 + 2];
            sure = sure && h > hmargin[k];
            maybe = maybe && h > -hmargin[k];
        }
        in_sure = sure;
        in_maybe = in_maybe || maybe;
    }
    in_maybe = in_maybe || in_sure;
    const bool both_sure = pos_sure && neg_sure, both_maybe = pos_maybe && neg_maybe;
    if (in_sure && !both_maybe) return 1;
    if (!in_maybe || both_sure) return 0;
    return 2;
}

// largest |height| of a base's own atoms off its own plane
__device__ __forceinline__ float st32_thickness(const SFrame &f, const float *q, uint32_t bm) {
    float t = 0.f;
#pragma unroll
    for (int a = 0; a < FR3D_BASE_SLOTS; ++a) {
        const float z = q[a * 3] * f.n0 + q[a * 3 + 1] * f.n1 + q[a * 3 + 2] * f.n2;
        if ((bm >> a) & 1u) t = fmaxf(t, fabsf(z));
    }
    return t;
}

__global__ void __launch_bounds__(ST32_WARPS * 32, 1)
classify_stack32_kernel(const float *__restrict__ rec, const int32_t *__restrict__ pair_i,
                        const int32_t *__restrict__ pair_j, const int32_t *__restrict__ pair_rank, int64_t pair_cap,
                        int32_t index_offset, fr3d_hit_t *__restrict__ hits, int64_t hit_cap, int64_t *counters,
                        const int32_t *__restrict__ worklist, int list_counter, int32_t *__restrict__ wl_unc,
                        unsigned long long *unc_count) {
    extern __shared__ __align__(16) float st32_smem[];
    float *shull = st32_smem;                               // [parents][7][3]
    float *smargin = shull + SCR_HULL;                      // [parents][7]
    float *sr2in = smargin + FR3D_N_PARENTS * 7;            // [parents], margin subtracted
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *rows = st32_smem + ST32_FIXED + warp * 32 * ST32_STRIDE;
    for (int item = threadIdx.x; item < SCR_HULL; item += blockDim.x) shull[item] = (float)(&c_tab.hull[0][0][0])[item];
    for (int item = threadIdx.x; item < FR3D_N_PARENTS * 7; item += blockDim.x) {
        const double *l = &c_tab.hull[0][0][0] + item * 3;
        smargin[item] = (float)(1e-3 * (fabs(l[0]) + fabs(l[1])) + 1e-5 * fabs(l[2]) + 1e-6);
    }
    if (threadIdx.x < FR3D_N_PARENTS) sr2in[threadIdx.x] = (float)c_tab.hull_r2in[threadIdx.x] - 1e-2f;
    __syncthreads();

    long long n_pairs = counters[list_counter];
    if (n_pairs > pair_cap) n_pairs = pair_cap;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long rounds = (n_pairs + stride - 1) / stride;
    int nx1 = 0, nx2 = 0, nxq = 0;
    if (first < n_pairs) {
        nxq = worklist[first];
        nx1 = pair_i[nxq]; nx2 = pair_j[nxq];
    }
    constexpr float INV_FX = (float)(1.0 / S32_FX);
#pragma unroll 1
    for (long long r = 0; r < rounds; ++r) {
        const long long p = first + r * stride;
        const bool active = p < n_pairs;
        const int n1 = nx1, n2 = nx2, q = nxq;
        nx1 = 0; nx2 = 0; nxq = 0;
        if (p + stride < n_pairs) {
            nxq = worklist[p + stride];
            nx1 = pair_i[nxq]; nx2 = pair_j[nxq];
        }
        if (!__any_sync(FULL, active)) continue;
        // stage the 32 nt2 records (80 words each; row k belongs to lane k)
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 20; ++j) {                      // 20 pieces per row: 4 lanes x 16 B on eight rows per instruction
            const int row = (j & 3) * 8 + (lane >> 2), piece = ((j >> 2) * 4 + (lane & 3)) * 4;
            const int nb = __shfl_sync(FULL, n2, row);
            cp_async16(rows + row * ST32_STRIDE + piece, rec + (size_t)nb * S32_REC + piece);
        }
        // nt1: read directly (the list keeps the pairs grouped by nt1, so these loads are near-broadcast)
        const float *A = rec + (size_t)n1 * S32_REC;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A + 32));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A + 64));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A + 79));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + (size_t)nx1 * S32_REC));
        const int c1x = __float_as_int(A[0]), c1y = __float_as_int(A[1]), c1z = __float_as_int(A[2]);
        SFrame f1;
        f1.a0 = A[3]; f1.a1 = A[4]; f1.a2 = A[5]; f1.n0 = A[6]; f1.n1 = A[7]; f1.n2 = A[8];
        const uint32_t bm1 = __float_as_uint(A[S32_MASKS]);
        const int par1 = (int)((bm1 >> 16) & 15u) - 1;
        cp_async_wait_all();
        __syncwarp();

        int result = -1;                 // -1 no label, -2 undecided in single precision, else the hit code
        if (active) {
            const float *Rs = rows + lane * ST32_STRIDE;     // this lane's row; its first 32 words go to registers
            float R[S32_BASE];
            {
                const float4 *R4 = reinterpret_cast<const float4 *>(Rs);
#pragma unroll
                for (int j = 0; j < S32_BASE / 4; ++j) {
                    const float4 t = R4[j];
                    R[4 * j] = t.x; R[4 * j + 1] = t.y; R[4 * j + 2] = t.z; R[4 * j + 3] = t.w;
                }
            }
            const uint32_t bm2 = __float_as_uint(R[S32_MASKS]);
            const int par2 = (int)((bm2 >> 16) & 15u) - 1;
            if (par1 >= 0 && par1 < FR3D_N_PARENTS && par2 >= 0 && par2 < FR3D_N_PARENTS) {
                SFrame f2;
                f2.a0 = R[3]; f2.a1 = R[4]; f2.a2 = R[5]; f2.n0 = R[6]; f2.n1 = R[7]; f2.n2 = R[8];
                const float nZ = f1.n0 * f2.n0 + f1.n1 * f2.n1 + f1.n2 * f2.n2;
                const float anZ = fabsf(nZ);
                if (((bm1 | bm2) >> 22) & 1u) {
                    result = -2;                                             // twin slots: fp64 path
                } else if (anZ < 0.5f - 1e-4f) {
                    result = -1;                                             // NA:1717
                } else if (anZ < 0.5f + 1e-4f) {
                    result = -2;
                } else {
                    const float e0 = (float)(__float_as_int(R[0]) - c1x) * INV_FX;
                    const float e1 = (float)(__float_as_int(R[1]) - c1y) * INV_FX;
                    const float e2 = (float)(__float_as_int(R[2]) - c1z) * INV_FX;
                    const float b10 = f1.n1 * f1.a2 - f1.n2 * f1.a1, b11 = f1.n2 * f1.a0 - f1.n0 * f1.a2, b12 = f1.n0 * f1.a1 - f1.n1 * f1.a0;
                    const float b20 = f2.n1 * f2.a2 - f2.n2 * f2.a1, b21 = f2.n2 * f2.a0 - f2.n0 * f2.a2, b22 = f2.n0 * f2.a1 - f2.n1 * f2.a0;
                    const float *Q1 = A + S32_BASE;
                    float Q2[FR3D_BASE_SLOTS * 3];              // nt2's base points in registers: twelve 128-bit loads
                    {
                        const float4 *R4 = reinterpret_cast<const float4 *>(Rs + S32_BASE);
#pragma unroll
                        for (int j = 0; j < FR3D_BASE_SLOTS * 3 / 4; ++j) {
                            const float4 t = R4[j];
                            Q2[4 * j] = t.x; Q2[4 * j + 1] = t.y; Q2[4 * j + 2] = t.z; Q2[4 * j + 3] = t.w;
                        }
                    }
                    float z21, z12;
                    int a21, a12;
                    const int o21 = st32_overlap(f1, b10, b11, b12, shull + par1 * 21, smargin + par1 * 7, sr2in[par1], Q2,
                                                 Rs + S32_BASE, e0, e1, e2, bm2, z21, a21);
                    const int o12 = st32_overlap(f2, b20, b21, b22, shull + par2 * 21, smargin + par2 * 7, sr2in[par2], Q1,
                                                 Q1, -e0, -e1, -e2, bm1, z12, a12);
                    if (o21 == 2 || o12 == 2) {
                        result = -2;
                    } else if (o21 == 1 || o12 == 1) {
                        const bool reverse = o12 == 1 && o21 != 1;
                        const float mvd = reverse ? z12 : z21;
                        int code = mvd > 0.f ? (nZ > 0.f ? 0 : 2) : (nZ > 0.f ? 1 : 3);
                        if (reverse && code < 2) code ^= 1;
                        const bool both = o21 == 1 && o12 == 1;
                        // minimum atom-atom distance (NA:1799-1829) against 1 and, for a true stack, 4
                        const float lb = fmaxf(fabsf(z21) - st32_thickness(f1, Q1, bm1), fabsf(z12) - st32_thickness(f2, Q2, bm2));
                        bool gt1 = lb > 1.f + ST32_DZ;
                        const bool true_possible = both && anZ > 0.6f - 1e-4f;
                        bool lt4 = false, undecided = false;
                        if (!gt1 || true_possible) {
                            // "mind < 4" is certain as soon as ONE atom pair is closer than 4 - margin, and with
                            // "mind > 1" certified from the heights the loop can stop there; it starts at the atom of
                            // nt1 closest to base 2's plane, which almost always has such a partner
                            const bool need_min = !gt1;
                            constexpr float NEAR2 = (4.f - ST32_DZ) * (4.f - ST32_DZ);
                            float m2 = 1e30f;
#pragma unroll 1
                            for (int i = 0; i < FR3D_BASE_SLOTS; ++i) {
                                const int v = (a12 + i) & (FR3D_BASE_SLOTS - 1);
                                if (!((bm1 >> v) & 1u)) continue;
                                if (!need_min && m2 < NEAR2) break;
                                const float a0 = Q1[v * 3] - e0, a1 = Q1[v * 3 + 1] - e1, a2 = Q1[v * 3 + 2] - e2;
#pragma unroll
                                for (int k = 0; k < FR3D_BASE_SLOTS; ++k) {
                                    const float d0 = Q2[k * 3] - a0, d1 = Q2[k * 3 + 1] - a1, d2 = Q2[k * 3 + 2] - a2;
                                    const float dd = d0 * d0 + d1 * d1 + d2 * d2;
                                    if (((bm2 >> k) & 1u) && dd < m2) m2 = dd;
                                }
                            }
                            if (!need_min && m2 < NEAR2) {
                                lt4 = true;                                      // (gt1 holds from the heights)
                            } else {                                             // the loop ran to its end: m2 is the minimum
                                const float m = m2 < 1e29f ? sqrtf(m2) : 9999.f;
                                if (fabsf(m - 1.f) <= ST32_DZ || (true_possible && fabsf(m - 4.f) <= ST32_DZ)) undecided = true;
                                gt1 = m > 1.f;
                                lt4 = m < 4.f;
                            }
                        }
                        if (true_possible && fabsf(anZ - 0.6f) <= 1e-4f) undecided = true;
                        if (undecided) result = -2;
                        else if (!gt1) result = -1;                              // NA:1707: mind < 1, no label
                        else result = code | ((true_possible && lt4) ? 0 : 4);   // NA:1710
                    }
                }
            }
        }
        // undecided pairs go to the fp64 kernel's list
        const unsigned bu = __ballot_sync(FULL, result == -2);
        if (bu) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(unc_count, (unsigned long long)__popc(bu));
            base = __shfl_sync(FULL, base, 0);
            if (result == -2) wl_unc[base + __popc(bu & ((1u << lane) - 1u))] = q;
        }
        // hits: one atomicAdd per warp
        const unsigned bh = __ballot_sync(FULL, result >= 0);
        if (bh) {
            long long base = 0;
            if (lane == 0) base = (long long)atomicAdd((unsigned long long *)&counters[1], (unsigned long long)__popc(bh));
            base = __shfl_sync(FULL, base, 0);
            const long long pos = base + __popc(bh & ((1u << lane) - 1u));
            if (result >= 0 && pos < hit_cap) {
                const int rank = pair_rank[q];
                const uint32_t w3 = (uint32_t)FR3D_SLOT_STACK | ((uint32_t)result << 8);
                uint4 *dst = reinterpret_cast<uint4 *>(hits + pos);
                dst[0] = make_uint4((uint32_t)(n1 + index_offset), (uint32_t)(n2 + index_offset),
                                    (uint32_t)(rank + 16 * index_offset), w3);
                dst[1] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
    }
}

}  // namespace fr3d
