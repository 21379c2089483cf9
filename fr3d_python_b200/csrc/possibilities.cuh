// K9: FR3D search, intersection of the pairwise constraint lists into complete possibilities.
//
// Replaces getPossibilities / extendFragment (fr3d/search/search.py:263-375) over buildSecondElementList (:377-398).
// The reference walks the fragments depth first, one Python set intersection per (fragment, later position).  The set
// of possibilities it returns is
//   { t : (t_a, t_b) in listOfPairs[a][b] for every a < b with a list ("full": t_a in universe[a]),
//         symbolic: all t_a distinct (:310);  geometric / mixed: the running distance error of every prefix
//         (t_0 .. t_m), m >= 2, is <= Q["cutoff"][m]  (:316-322) }
// (the look-ahead tests of :282-292 only prune fragments that cannot complete).  Here the fragments advance one position
// per launch, breadth first: one WARP per fragment, the lanes take the candidates of the first constrained list
// ("driver") and test them against the other lists by binary search in CSR rows; count -> scan -> emit keeps the
// fragments in lexicographic order, so the result is the reference's set, sorted.  The distance error is accumulated in
// the reference's order, including its quirks: the error against the LAST element of the fragment is not added
// (`for j in range(0, n)`, :319), the first pair's error is carried but never tested (:358-366).
#pragma once

#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/fr3d_b200.h"

namespace fr3d {

constexpr int POSS_WARPS = 8;

__device__ __forceinline__ bool csr_contains(const int32_t *cols, int lo, int hi, int c) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int v = cols[mid];
        if (v == c) return true;
        if (v < c) lo = mid + 1; else hi = mid;
    }
    return false;
}

// np.linalg.norm(c_a - c_b) (search.py:133): sqrt of the 3-term dot product, which numpy's BLAS accumulates with fused
// multiply-adds in index order (checked against numpy in tests/test_possibilities.py)
__device__ __forceinline__ double unit_distance(const double *centers, int a, int b) {
    const double x = centers[(size_t)a * 3] - centers[(size_t)b * 3];
    const double y = centers[(size_t)a * 3 + 1] - centers[(size_t)b * 3 + 1];
    const double z = centers[(size_t)a * 3 + 2] - centers[(size_t)b * 3 + 2];
    return sqrt(__fma_rn(z, z, __fma_rn(y, y, __dmul_rn(x, x))));
}

// level = index of the fragment's last position (fragments have level + 1 elements); the new position is p = level + 1.
// EMIT = false: counts[f] = number of extensions of fragment f.  EMIT = true: offsets[f] = exclusive prefix of the counts;
// the extended fragments are written at offsets[f] .. in ascending order of the new element.
template <bool EMIT>
__global__ void __launch_bounds__(POSS_WARPS * 32)
poss_extend_kernel(const fr3d_search_query_t *__restrict__ qp, int level, const int32_t *__restrict__ frags,
                   const double *__restrict__ errs, long long n_frags, long long *__restrict__ counts,
                   int32_t *__restrict__ out_frags, double *__restrict__ out_errs, long long out_cap, int32_t *flags) {
    constexpr unsigned ALL = 0xFFFFFFFFu;
    const fr3d_search_query_t &q = *qp;
    const int lane = threadIdx.x & 31;
    const int len = level + 1, p = level + 1, K = FR3D_SEARCH_MAX_POSITIONS;
    if (EMIT && counts[n_frags] > out_cap) return;            // capacity: the caller re-runs with a larger buffer
    // the new element must be one the reference would accept as "currentFragment[-1] in secondElementList[p][i]"
    // for a later full list: a member of universe[p]
    bool need_universe = false;
    for (int b = p + 1; b < q.n_positions; ++b) need_universe |= q.rowptr[p * K + b] == nullptr;
    int driver = -1;
    for (int j = 0; j <= level && driver < 0; ++j)
        if (q.rowptr[j * K + p]) driver = j;
    const long long warps = (long long)gridDim.x * POSS_WARPS;
    for (long long f = (long long)blockIdx.x * POSS_WARPS + (threadIdx.x >> 5); f < n_frags; f += warps) {
        const int mine = lane < len ? frags[f * len + lane] : -1;
        if (driver < 0) {                                     // "next position to be added is a full list" (:295-300)
            if (lane == 0) { atomicOr(flags, 1); if (!EMIT) counts[f] = 0; }
            continue;
        }
        const int td = __shfl_sync(ALL, mine, driver);
        const int32_t *dcols = q.cols[driver * K + p];
        const int beg = q.rowptr[driver * K + p][td], end = q.rowptr[driver * K + p][td + 1];
        const double prev = q.symbolic ? 0.0 : errs[f];
        long long pos = EMIT ? counts[f] : 0;
        int total = 0;
        for (int c0 = beg; c0 < end; c0 += 32) {
            const int ci = c0 + lane;
            bool ok = ci < end;
            const int c = ok ? dcols[ci] : -1;
            double err = prev;
            if (ok && need_universe && !(q.universe[p] && q.universe[p][c])) ok = false;
            for (int j = 0; j <= level; ++j) {                // warp-uniform loop; shuffles need every lane
                const int tj = __shfl_sync(ALL, mine, j);
                if (!ok) continue;
                if (j != driver && q.rowptr[j * K + p]) {
                    const int32_t *rp = q.rowptr[j * K + p];
                    if (!csr_contains(q.cols[j * K + p], rp[tj], rp[tj + 1], c)) ok = false;
                }
                if (q.symbolic) { if (tj == c) ok = false; }                       // no nucleotide twice (:310)
                else if (j < level) {                                              // range(0, n): the last one is skipped
                    const double e = q.perm_dist[j * K + p] - unit_distance(q.centers, tj, c);
                    err = __dadd_rn(err, __dmul_rn(e, e));
                }
            }
            if (ok && !q.symbolic && !(err <= q.cutoff[p])) ok = false;            // :322
            const unsigned b = __ballot_sync(ALL, ok);
            if (EMIT && ok) {
                const long long at = pos + __popc(b & ((1u << lane) - 1u));
                int32_t *dst = out_frags + at * (len + 1);
                for (int j = 0; j < len; ++j) dst[j] = frags[f * len + j];
                dst[len] = c;
                if (!q.symbolic) out_errs[at] = err;
            }
            pos += __popc(b);
            total += __popc(b);
        }
        if (!EMIT && lane == 0) counts[f] = total;
    }
}

// exclusive scan of counts[0 .. n) in place, the total in counts[n]; one CTA
__global__ void __launch_bounds__(1024) poss_scan_kernel(long long *counts, long long n) {
    __shared__ long long s_warp[32];
    __shared__ long long s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (long long base = 0; base < n; base += 1024) {
        const long long i = base + threadIdx.x;
        const long long v = i < n ? counts[i] : 0;
        long long incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const long long t = __shfl_up_sync(0xFFFFFFFFu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const long long w = s_warp[lane];
            long long wi = w;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const long long t = __shfl_up_sync(0xFFFFFFFFu, wi, off);
                if (lane >= off) wi += t;
            }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        const long long carry = s_carry;
        if (i < n) counts[i] = carry + s_warp[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[31] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[n] = s_carry;
}

// The first pairs (search.py:351-366): every pair of listOfPairs[0][1] whose first element is acceptable to every later
// list; geometric / mixed: its distance error is carried, not tested.
__global__ void poss_first_pairs_kernel(const fr3d_search_query_t *__restrict__ qp, const int32_t *__restrict__ pairs,
                                        long long n_pairs, long long *__restrict__ counts) {
    const fr3d_search_query_t &q = *qp;
    const int K = FR3D_SEARCH_MAX_POSITIONS;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (long long)gridDim.x * blockDim.x) {
        const int a = pairs[i * 2], b = pairs[i * 2 + 1];
        bool ok = true;
        for (int j = 1; j < q.n_positions && ok; ++j) {       // firstPair[0] in secondElementList[0][j]
            const int32_t *rp = q.rowptr[j];                  // row 0 of the K x K table
            if (rp) ok = rp[a + 1] > rp[a];
            else ok = q.universe[0] && q.universe[0][a];
        }
        for (int j = 2; j < q.n_positions && ok; ++j)         // firstPair[1] in secondElementList[1][j] (:282, n = 1)
            if (!q.rowptr[K + j]) ok = q.universe[1] && q.universe[1][b];
        counts[i] = ok ? 1 : 0;
    }
}

__global__ void poss_first_emit_kernel(const fr3d_search_query_t *__restrict__ qp, const int32_t *__restrict__ pairs,
                                       long long n_pairs, const long long *__restrict__ offsets,
                                       int32_t *__restrict__ out_frags, double *__restrict__ out_errs, long long out_cap) {
    const fr3d_search_query_t &q = *qp;
    if (offsets[n_pairs] > out_cap) return;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (long long)gridDim.x * blockDim.x) {
        if (offsets[i + 1] == offsets[i]) continue;
        const long long at = offsets[i];
        const int a = pairs[i * 2], b = pairs[i * 2 + 1];
        out_frags[at * 2] = a;
        out_frags[at * 2 + 1] = b;
        if (!q.symbolic) {
            const double e = q.perm_dist[1] - unit_distance(q.centers, a, b);
            out_errs[at] = __dmul_rn(e, e);
        }
    }
}

}  // namespace fr3d
