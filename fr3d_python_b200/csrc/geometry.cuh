// K5 batched Kabsch superposition and K6 all-vs-all / one-vs-many geometric discrepancy.
//
// Replaces besttransformation(_weighted) (fr3d/geometry/superpositions.py:10-121),
// matrix_discrepancy / matrix_discrepancy_cutoff (fr3d/geometry/discrepancy.py:165-313;
// same arithmetic as fr3d/search/discrepancy.py:165-313 with default weights) and
// angle_of_rotation (fr3d/geometry/angleofrotation.py:4-7).
//
// K6 is FP64-pipe bound: one thread per (i, j) instance pair, a 32 x 32 tile of
// pairs per CTA with the 64 instances' centred coordinates and rotations staged
// in shared memory (each is reused 32 times).  Flops per pair ~ 154 n + Jacobi.
#pragma once

#include "rotfit.cuh"

namespace fr3d {

// ---------------------------------------------------------------- K5 -------
// One thread per problem (ragged offsets).  A thread walking "its" rows in global memory makes every load of the warp
// touch 32 different sectors (2.0 TB/s of DRAM traffic at 31 % of the HBM peak, profiles/geometry_r2_raw_metrics.txt), so
// the rows of a warp's 32 problems - one contiguous block of set1 / set2 (/ w) - are first copied into shared memory with
// coalesced cp.async, the three passes (means, covariance, fitted coordinates + sse) run out of shared memory, and the
// fitted coordinates are written back to global memory the same coalesced way.  A warp whose 32 problems have more than
// KABSCH_ROWS rows in total (long point sets) works from global memory as before.
constexpr int KABSCH_WARPS = 4;
#ifndef FR3D_KABSCH_ROWS
#define FR3D_KABSCH_ROWS 352
#endif
constexpr int KABSCH_ROWS = FR3D_KABSCH_ROWS;    // rows staged per warp (352: 11 points per problem on average; 0: no staging)
constexpr size_t KABSCH_SMEM = sizeof(double) * KABSCH_WARPS * KABSCH_ROWS * 7;      // with weights; 6 / 7 of it without

__global__ void __launch_bounds__(KABSCH_WARPS * 32)
kabsch_batched_kernel(const double *__restrict__ set1, const double *__restrict__ set2, const double *__restrict__ w,
                      const int32_t *__restrict__ off, int n_prob, double *__restrict__ Uo, double *__restrict__ mean1o,
                      double *__restrict__ mean2o, double *__restrict__ rmsdo, double *__restrict__ sseo,
                      double *__restrict__ new1o) {
    extern __shared__ double kb_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p0 = (blockIdx.x * KABSCH_WARPS + warp) * 32;
    if (p0 >= n_prob) return;
    const int p = p0 + lane;
    const int R0 = off[p0], R1 = off[min(p0 + 32, n_prob)];
    const int R = R1 - R0;
    const bool staged = R <= KABSCH_ROWS;                                  // warp uniform
    double *s1 = kb_smem + (size_t)warp * KABSCH_ROWS * (w ? 7 : 6), *s2 = s1 + KABSCH_ROWS * 3, *sw = s2 + KABSCH_ROWS * 3;
    if (staged) {
        for (int t = lane; t < R * 3; t += 32) {
            cp_async8(s1 + t, set1 + (size_t)R0 * 3 + t);
            cp_async8(s2 + t, set2 + (size_t)R0 * 3 + t);
        }
        if (w) for (int t = lane; t < R; t += 32) cp_async8(sw + t, w + R0 + t);
        cp_async_wait_all();
        __syncwarp();
    }
    // a1[k * 3 + c] / a2 / aw[k] with GLOBAL row numbers k, whichever memory holds them; the fitted coordinates of a staged
    // warp go to its set1 rows in place
    const double *a1 = staged ? s1 - (size_t)R0 * 3 : set1;
    const double *a2 = staged ? s2 - (size_t)R0 * 3 : set2;
    const double *aw = w ? (staged ? sw - R0 : w) : nullptr;
    double *out1 = staged ? s1 - (size_t)R0 * 3 : new1o;
    if (p < n_prob) {
        const int r0 = off[p], r1 = off[p + 1];
        const int len = r1 - r0;
        double m1[3] = {0, 0, 0}, m2[3] = {0, 0, 0};
        for (int k = r0; k < r1; ++k) {
#pragma unroll
            for (int c = 0; c < 3; ++c) { m1[c] += a1[(size_t)k * 3 + c]; m2[c] += a2[(size_t)k * 3 + c]; }
        }
        const double L = (double)len;
#pragma unroll
        for (int c = 0; c < 3; ++c) { m1[c] /= L; m2[c] /= L; }
        // M = dev1^T diag(w) dev2   (= A^T of superpositions.py:53 / :109)
        double M[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        double ssq = 0.0;                                      // sum |a|^2 + sum |w b|^2: twice the eigenvalue bound
        for (int k = r0; k < r1; ++k) {
            const double wk = aw ? aw[k] : 1.0;
            const double x0 = a1[(size_t)k * 3 + 0] - m1[0], x1 = a1[(size_t)k * 3 + 1] - m1[1],
                         x2 = a1[(size_t)k * 3 + 2] - m1[2];
            const double b0 = (a2[(size_t)k * 3 + 0] - m2[0]) * wk, b1 = (a2[(size_t)k * 3 + 1] - m2[1]) * wk,
                         b2 = (a2[(size_t)k * 3 + 2] - m2[2]) * wk;
            M[0][0] += x0 * b0; M[0][1] += x0 * b1; M[0][2] += x0 * b2;
            M[1][0] += x1 * b0; M[1][1] += x1 * b1; M[1][2] += x1 * b2;
            M[2][0] += x2 * b0; M[2][1] += x2 * b1; M[2][2] += x2 * b2;
            ssq += (x0 * x0 + x1 * x1 + x2 * x2) + (b0 * b0 + b1 * b1 + b2 * b2);
        }
        double Q[3][3];
        // largest eigenvector of Horn's matrix from its characteristic quartic + a pivoted adjugate column (rotfit.cuh; Jacobi
        // fallback when the optimum is not unique): the cyclic Jacobi sweeps alone kept this kernel at 0.43 ms whatever the
        // memory access pattern was
        horn_rotation_pivot(M, 0.5 * ssq, Q);
        // U = Q^T (row-vector convention: new1 = dev1 . U)
        double sse = 0.0;
        for (int k = r0; k < r1; ++k) {
            const double x0 = a1[(size_t)k * 3 + 0] - m1[0], x1 = a1[(size_t)k * 3 + 1] - m1[1],
                         x2 = a1[(size_t)k * 3 + 2] - m1[2];
            const double n0 = x0 * Q[0][0] + x1 * Q[0][1] + x2 * Q[0][2];
            const double n1 = x0 * Q[1][0] + x1 * Q[1][1] + x2 * Q[1][2];
            const double n2 = x0 * Q[2][0] + x1 * Q[2][1] + x2 * Q[2][2];
            if (new1o) { out1[(size_t)k * 3 + 0] = n0; out1[(size_t)k * 3 + 1] = n1; out1[(size_t)k * 3 + 2] = n2; }
            const double e0 = n0 - (a2[(size_t)k * 3 + 0] - m2[0]), e1 = n1 - (a2[(size_t)k * 3 + 1] - m2[1]),
                         e2 = n2 - (a2[(size_t)k * 3 + 2] - m2[2]);
            sse += e0 * e0 + e1 * e1 + e2 * e2;               // unweighted, RMSD.py:27
        }
        if (Uo) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) Uo[(size_t)p * 9 + r * 3 + c] = Q[c][r];
        }
        if (mean1o) { mean1o[(size_t)p * 3 + 0] = m1[0]; mean1o[(size_t)p * 3 + 1] = m1[1]; mean1o[(size_t)p * 3 + 2] = m1[2]; }
        if (mean2o) { mean2o[(size_t)p * 3 + 0] = m2[0]; mean2o[(size_t)p * 3 + 1] = m2[1]; mean2o[(size_t)p * 3 + 2] = m2[2]; }
        if (sseo) sseo[p] = sse;
        if (rmsdo) rmsdo[p] = sqrt(sse / L);                    // RMSD.py:15
    }
    if (staged && new1o) {
        __syncwarp();
        for (int t = lane; t < R * 3; t += 32) new1o[(size_t)R0 * 3 + t] = s1[t];
    }
}

// ---------------------------------------------------------------- K6 -------
// Discrepancy of instance A (centred coords ca[n][3], rotations ra[n][9]) and B, n > 2
// (discrepancy.py:192-204).  `cutoff2 >= 0` enables the early exit of
// matrix_discrepancy_cutoff (discrepancy.py:263-281) -> NaN.
template <int MAXN>
__device__ __forceinline__ double discrepancy_pair(const double *ca, const double *ra, const uint8_t *ha,
                                                   const double *cb, const double *rb, const uint8_t *hb, int n,
                                                   double angle_weight, const double *cw, double cutoff2) {
    // A = dev2^T diag(w) dev1 with set1 = A's centres, set2 = B's; M = A^T
    double M[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    double ga = 0.0, gb = 0.0;
    for (int k = 0; k < n; ++k) {
        const double wk = cw ? cw[k] : 1.0;
        const double a0 = ca[k * 3 + 0], a1 = ca[k * 3 + 1], a2 = ca[k * 3 + 2];
        const double b0 = cb[k * 3 + 0] * wk, b1 = cb[k * 3 + 1] * wk, b2 = cb[k * 3 + 2] * wk;
        M[0][0] += a0 * b0; M[0][1] += a0 * b1; M[0][2] += a0 * b2;
        M[1][0] += a1 * b0; M[1][1] += a1 * b1; M[1][2] += a1 * b2;
        M[2][0] += a2 * b0; M[2][1] += a2 * b1; M[2][2] += a2 * b2;
        ga += a0 * a0 + a1 * a1 + a2 * a2;
        gb += b0 * b0 + b1 * b1 + b2 * b2;
    }
    double Q[3][3];
    horn_rotation_fast(M, 0.5 * (ga + gb), Q);          // U = Q^T
    double sse = 0.0;
    for (int k = 0; k < n; ++k) {
        const double a0 = ca[k * 3 + 0], a1 = ca[k * 3 + 1], a2 = ca[k * 3 + 2];
        const double e0 = a0 * Q[0][0] + a1 * Q[0][1] + a2 * Q[0][2] - cb[k * 3 + 0];
        const double e1 = a0 * Q[1][0] + a1 * Q[1][1] + a2 * Q[1][2] - cb[k * 3 + 1];
        const double e2 = a0 * Q[2][0] + a1 * Q[2][1] + a2 * Q[2][2] - cb[k * 3 + 2];
        sse += e0 * e0 + e1 * e1 + e2 * e2;
    }
    double total = sse;
    if (cutoff2 >= 0 && total > cutoff2) return nan("");
    double orient = 0.0;
    for (int k = 0; k < n; ++k) {
        if ((ha == nullptr || ha[k]) && (hb == nullptr || hb[k])) {
            // trace(U . r2 . r1^T) with r1 = A's rotation, r2 = B's:  sum_ij (U r2)_ij r1_ij, U = Q^T
            const double *r1 = ra + k * 9, *r2 = rb + k * 9;
            double tr = 0.0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const double ur2 = Q[0][i] * r2[0 * 3 + j] + Q[1][i] * r2[1 * 3 + j] + Q[2][i] * r2[2 * 3 + j];
                    tr += ur2 * r1[i * 3 + j];
                }
            }
            const double ang = angle_from_trace(tr);
            if (cutoff2 >= 0) {
                total += ang * ang;
                if (total > cutoff2) return nan("");
            } else {
                orient += ang * ang;
            }
        }
    }
    if (cutoff2 >= 0) return sqrt(total) / (double)n;
    return sqrt(sse + angle_weight * orient) / (double)n;
}

// n == 2 closed form (discrepancy.py:206-231).  ca/cb are RAW centres here.
__device__ __forceinline__ double discrepancy_pair2(const double *c1, const double *r1, const double *c2,
                                                    const double *r2, double angle_weight) {
    // R1 = r1[1]^T r1[0];  R2 = r2[0]^T r2[1]
    double R1[9], R2[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            R1[i * 3 + j] = r1[9 + 0 * 3 + i] * r1[0 * 3 + j] + r1[9 + 1 * 3 + i] * r1[1 * 3 + j] + r1[9 + 2 * 3 + i] * r1[2 * 3 + j];
            R2[i * 3 + j] = r2[0 * 3 + i] * r2[9 + 0 * 3 + j] + r2[1 * 3 + i] * r2[9 + 1 * 3 + j] + r2[2 * 3 + i] * r2[9 + 2 * 3 + j];
        }
    // ang1 = angle(R1 R2): trace = sum_ij R1_ij R2_ji ; ang2 = angle(R1^T R2^T): trace = sum_ij R1_ji R2_ij (same)
    double tr = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) tr += R1[i * 3 + j] * R2[j * 3 + i];
    const double ang1 = angle_from_trace(tr);
    const double ang2 = ang1;
    double D1[3], D2[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double T1 = 0, T2 = 0, S1 = 0, S2 = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            T1 += (c1[3 + k] - c1[k]) * r1[k * 3 + j];
            T2 += (c1[k] - c1[3 + k]) * r1[9 + k * 3 + j];
            S1 += (c2[3 + k] - c2[k]) * r2[k * 3 + j];
            S2 += (c2[k] - c2[3 + k]) * r2[9 + k * 3 + j];
        }
        D1[j] = T1 - S1;
        D2[j] = T2 - S2;
    }
    double d = sqrt(D1[0] * D1[0] + D1[1] * D1[1] + D1[2] * D1[2] + (angle_weight * ang1) * (angle_weight * ang1));
    d += sqrt(D2[0] * D2[0] + D2[1] * D2[1] + D2[2] * D2[2] + (angle_weight * ang2) * (angle_weight * ang2));
    return d * 0.17677669529663687;
}

static constexpr int DISC_TILE = 32;
static constexpr int DISC_MAXN = 16;      // positions per instance supported by the tiled kernel

// dynamic shared memory of discrepancy_all_vs_all_kernel for n positions (the opt-in in fr3d_init uses n = DISC_MAXN)
constexpr size_t disc_smem_bytes(int n) {
    return (size_t)DISC_TILE * (size_t)n * 12 * 2 * sizeof(double) + 2 * DISC_TILE * DISC_MAXN +
           (size_t)DISC_TILE * (DISC_TILE + 1) * sizeof(double);
}

// grid: (ceil((N - col0)/32), ceil(rows/32)); block 32x8 threads, each thread 4 rows of the tile.
// Writes out[(r - row0) * ld + (c - col0)] for row0 <= r < row1, col0 <= c < N (col0 a multiple of 32).  col0 = 0,
// ld = N: whole rows; col0 = row0, ld = N - row0: only the part of a row block at or right of the diagonal (the
// multi-GPU split of the upper triangle, parallel.triangle_blocks; the mirror is fr3d_discrepancy_assemble).
#ifndef FR3D_DISC_MINB
#define FR3D_DISC_MINB 3      // 2 (116 registers) 21.83 ms, 3 (80, no spills) 20.12, 4 (64, spills) 19.96 for 20 000 x 20 000, n = 7
#endif
__global__ void __launch_bounds__(256, FR3D_DISC_MINB)
discrepancy_all_vs_all_kernel(const double *__restrict__ centers, const double *__restrict__ rots,
                              const uint8_t *__restrict__ has_rot, int N, int n, double angle_weight, int row0,
                              int row1, int col0, int ld, double *__restrict__ out) {
    extern __shared__ double smem[];
    // layout: rowc[32][n*3] rowr[32][n*9] colc[32][n*3] colr[32][n*9] + flags
    const int cs = n * 3, rs = n * 9;
    double *rowc = smem;
    double *rowr = rowc + DISC_TILE * cs;
    double *colc = rowr + DISC_TILE * rs;
    double *colr = colc + DISC_TILE * cs;
    uint8_t *rowh = reinterpret_cast<uint8_t *>(colr + DISC_TILE * rs);
    uint8_t *colh = rowh + DISC_TILE * DISC_MAXN;
    double *tile = reinterpret_cast<double *>(colh + DISC_TILE * DISC_MAXN);     // [32][33] mirrored results

    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int rbase = row0 + blockIdx.y * DISC_TILE;
    const int cbase = col0 + blockIdx.x * DISC_TILE;
    // full-matrix launches compute only tiles on or above the diagonal and mirror them: the reference
    // fills both triangles from one evaluation as well (FR3D.py:439-441)
    const bool mirror = (row0 == 0 && row1 == N && col0 == 0);
    if (mirror && blockIdx.x < blockIdx.y) return;

    // stage: raw centres, rotations (coalesced: consecutive threads read consecutive doubles)
    for (int e = tid; e < DISC_TILE * cs; e += 256) {
        const int t = e / cs, r = rbase + t, c = cbase + t;
        rowc[e] = (r < row1) ? centers[(size_t)r * cs + (e - t * cs)] : 0.0;
        colc[e] = (c < N) ? centers[(size_t)c * cs + (e - t * cs)] : 0.0;
    }
    for (int e = tid; e < DISC_TILE * rs; e += 256) {
        const int t = e / rs, r = rbase + t, c = cbase + t;
        rowr[e] = (r < row1) ? rots[(size_t)r * rs + (e - t * rs)] : 0.0;
        colr[e] = (c < N) ? rots[(size_t)c * rs + (e - t * rs)] : 0.0;
    }
    for (int e = tid; e < DISC_TILE * n; e += 256) {
        const int t = e / n, k = e - t * n, r = rbase + t, c = cbase + t;
        rowh[t * DISC_MAXN + k] = (has_rot && r < row1) ? has_rot[(size_t)r * n + k] : 1;
        colh[t * DISC_MAXN + k] = (has_rot && c < N) ? has_rot[(size_t)c * n + k] : 1;
    }
    __syncthreads();
    if (n > 2) {
        // centre each staged instance in place (superpositions.py:105-108); one thread per instance
        if (tid < 2 * DISC_TILE) {
            double *cc = (tid < DISC_TILE) ? rowc + tid * cs : colc + (tid - DISC_TILE) * cs;
            double m0 = 0, m1 = 0, m2 = 0;
            for (int k = 0; k < n; ++k) { m0 += cc[k * 3]; m1 += cc[k * 3 + 1]; m2 += cc[k * 3 + 2]; }
            const double L = (double)n;
            m0 /= L; m1 /= L; m2 /= L;
            for (int k = 0; k < n; ++k) { cc[k * 3] -= m0; cc[k * 3 + 1] -= m1; cc[k * 3 + 2] -= m2; }
        }
        __syncthreads();
    }
    const int c = cbase + threadIdx.x;
#pragma unroll 1
    for (int rr = threadIdx.y; rr < DISC_TILE; rr += 8) {
        const int r = rbase + rr;
        if (r >= row1 || c >= N) continue;
        double d = 0.0;
        if (r != c) {
            // the reference computes (i, j) with i < j and mirrors it (FR3D.py:433-442): keep that
            // argument order so both triangles hold the same bits
            const bool swap = r > c;
            const double *ca = swap ? colc + threadIdx.x * cs : rowc + rr * cs;
            const double *ra = swap ? colr + threadIdx.x * rs : rowr + rr * rs;
            const uint8_t *ha = swap ? colh + threadIdx.x * DISC_MAXN : rowh + rr * DISC_MAXN;
            const double *cb = swap ? rowc + rr * cs : colc + threadIdx.x * cs;
            const double *rb = swap ? rowr + rr * rs : colr + threadIdx.x * rs;
            const uint8_t *hb = swap ? rowh + rr * DISC_MAXN : colh + threadIdx.x * DISC_MAXN;
            if (n > 2) d = discrepancy_pair<DISC_MAXN>(ca, ra, ha, cb, rb, hb, n, angle_weight, nullptr, -1.0);
            else d = discrepancy_pair2(ca, ra, cb, rb, angle_weight);
        }
        out[(size_t)(r - row0) * ld + (c - col0)] = d;
        if (mirror && blockIdx.x > blockIdx.y) tile[threadIdx.x * (DISC_TILE + 1) + rr] = d;
    }
    if (mirror && blockIdx.x > blockIdx.y) {
        // transposed tile: rows c, columns r, written with consecutive threads along r (coalesced)
        __syncthreads();
#pragma unroll 1
        for (int cc = threadIdx.y; cc < DISC_TILE; cc += 8) {
            const int cg = cbase + cc, rg = rbase + threadIdx.x;
            if (cg < N && rg < N) out[(size_t)cg * N + rg] = tile[cc * (DISC_TILE + 1) + threadIdx.x];
        }
    }
}

// Place one upper row block blk [(r1 - r0)][N - r0] (the output of a col0 = row0 launch) into the N x N matrix D and
// mirror it: D[r][c] = D[c][r] = blk[r - r0][c - r0] for r0 <= r < r1, c >= r0 (FR3D.py:439-441).  32 x 32 tiles
// through shared memory so that both the row-wise and the transposed writes are coalesced.
__global__ void __launch_bounds__(256)
discrepancy_assemble_kernel(const double *__restrict__ blk, int N, int r0, int r1, double *__restrict__ D) {
    __shared__ double tile[DISC_TILE][DISC_TILE + 1];
    const int cbase = r0 + blockIdx.x * DISC_TILE, rbase = r0 + blockIdx.y * DISC_TILE;
    const int ld = N - r0;
    for (int rr = threadIdx.y; rr < DISC_TILE; rr += 8) {
        const int r = rbase + rr, c = cbase + threadIdx.x;
        double v = 0.0;
        if (r < r1 && c < N) {
            v = blk[(size_t)(r - r0) * ld + (c - r0)];
            D[(size_t)r * N + c] = v;
        }
        tile[rr][threadIdx.x] = v;
    }
    __syncthreads();
    if (cbase >= r1) {                       // right of the block's own diagonal square (which holds both triangles)
        for (int cc = threadIdx.y; cc < DISC_TILE; cc += 8) {
            const int c = cbase + cc, r = rbase + threadIdx.x;
            if (c < N && r < r1) D[(size_t)c * N + r] = tile[threadIdx.x][cc];
        }
    }
}

// FP64 peak probe: 8 independent fused multiply-add chains per thread, `iters` rounds; 2 * 8 * iters flops per thread.
// Timed by the caller (bench.py) to get the DFMA throughput the discrepancy roofline is quoted against.
__global__ void __launch_bounds__(256) fp64_fma_probe_kernel(long long iters, double seed, double *__restrict__ sink) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, b = 1e-9;
    for (long long k = 0; k < iters; ++k) {
        a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
        a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
    }
    const double total = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (total == 123.456) sink[0] = total;          // never true: keeps the chains alive
}

// one query against M candidates, one thread per candidate (search.py:1073-1085)
__global__ void __launch_bounds__(128)
discrepancy_one_vs_many_kernel(const double *__restrict__ qc, const double *__restrict__ qr,
                               const uint8_t *__restrict__ qh, const double *__restrict__ centers,
                               const double *__restrict__ rots, const uint8_t *__restrict__ has_rot, int M, int n,
                               double cutoff, double angle_weight, const double *__restrict__ cw,
                               double *__restrict__ out) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const double *cb_raw = centers + (size_t)m * n * 3;
    const double *rb = rots + (size_t)m * n * 9;
    const uint8_t *hb = has_rot ? has_rot + (size_t)m * n : nullptr;
    if (n == 2) {
        out[m] = discrepancy_pair2(qc, qr, cb_raw, rb, angle_weight);
        return;
    }
    double ca[DISC_MAXN * 3], cb[DISC_MAXN * 3];
    double ma[3] = {0, 0, 0}, mb[3] = {0, 0, 0};
    for (int k = 0; k < n; ++k)
        for (int c = 0; c < 3; ++c) { ma[c] += qc[k * 3 + c]; mb[c] += cb_raw[k * 3 + c]; }
    const double L = (double)n;
    for (int c = 0; c < 3; ++c) { ma[c] /= L; mb[c] /= L; }
    for (int k = 0; k < n; ++k)
        for (int c = 0; c < 3; ++c) { ca[k * 3 + c] = qc[k * 3 + c] - ma[c]; cb[k * 3 + c] = cb_raw[k * 3 + c] - mb[c]; }
    const double cutoff2 = (cutoff >= 0) ? ((double)n * cutoff) * ((double)n * cutoff) : -1.0;
    out[m] = discrepancy_pair<DISC_MAXN>(ca, qr, qh, cb, rb, hb, n, angle_weight, cw, cutoff2);
}

}  // namespace fr3d
