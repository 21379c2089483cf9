// K8: ordering by similarity of the all-vs-all discrepancy matrix (last part).
//
// Replaces treePenalizedPathLength and its parts in BOTH live copies of the reference:
//   fr3d/ordering/orderBySimilarity.py  :10-41 treePenalty, :43-64 normalizePenaltyMatrix, :66-80 twoOptSwap,
//                                       :82-115 greedyInsertionPathLength, :131-163 the repetitions, :165-186, :237-245
//   fr3d/search/orderBySimilarity.py    :9-42, :49-103 (penalty x 5, greedy insertion only; search/FR3D.py:463)
// on a matrix that stays in HBM (the K6 output), so the n x n matrix never crosses PCIe for the ordering.
//
// What is parallel and what is not.  The repetitions are independent: one CTA per repetition.  Inside a repetition
// the insertions (and the 2-opt swaps) are sequential by definition, but every insertion scores all positions of
// the current path at once (one lexicographic (score, position-priority) minimum per step reproduces the
// reference's first-strictly-smaller scan), and the 2-opt scan looks for the first improving j of a row in one sweep.
// Average linkage (scipy's nearest-neighbour chain for method "average"; the reference only uses the merge heights)
// runs in ONE CTA: ~3n chain steps, each a row-wide argmin.  All arithmetic that decides an order is done with
// explicit round-to-nearest intrinsics in the reference's own operation order - no FMA contraction - so scores,
// penalties and therefore the orders are bit-identical to the reference's floats.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>

namespace fr3d {

constexpr int ORD_THREADS = 1024;

// lexicographic minimum of (v, k) over the block; every thread gets the result.  s_v / s_k: 32 entries each.
__device__ __forceinline__ void block_argmin(double &v, int &k, double *s_v, int *s_k) {
    constexpr unsigned ALL = 0xFFFFFFFFu;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const double ov = __shfl_xor_sync(ALL, v, off);
        const int ok = __shfl_xor_sync(ALL, k, off);
        if (ov < v || (ov == v && ok < k)) { v = ov; k = ok; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();                       // the previous round's readers are done with s_v / s_k
    if (lane == 0) { s_v[warp] = v; s_k[warp] = k; }
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = lane < nw ? s_v[lane] : __longlong_as_double(0x7FF0000000000000ll);
    k = lane < nw ? s_k[lane] : 0x7FFFFFFF;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const double ov = __shfl_xor_sync(ALL, v, off);
        const int ok = __shfl_xor_sync(ALL, k, off);
        if (ov < v || (ov == v && ok < k)) { v = ov; k = ok; }
    }
}

__device__ __forceinline__ int block_max(int k, int *s_k) {
    constexpr unsigned ALL = 0xFFFFFFFFu;
    k = __reduce_max_sync(ALL, k);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_k[warp] = k;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    k = lane < nw ? s_k[lane] : (int)0x80000000;
    return __reduce_max_sync(ALL, k);
}

// ------------------------------------------------------------------------------------------------------------
// linkage(squareform(distance), "average"): scipy refuses a matrix that is not symmetric, has a non-zero diagonal or
// a non-finite entry; flags: 1 asymmetric, 2 diagonal, 4 non-finite.  The same pass makes the working copy.
__global__ void linkage_prepare_kernel(const double *__restrict__ d, int n, double *__restrict__ W, int32_t *flags) {
    const size_t total = (size_t)n * n;
    int f = 0;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / n), j = (int)(t % n);
        const double v = d[t];
        W[t] = v;
        if (i == j) { if (v != 0.0) f |= 2; }
        else {
            if (!(v == d[(size_t)j * n + i])) f |= 1;
            if (!(fabs(v) <= 1.79769313486231570e308)) f |= 4;
        }
    }
    if (f) atomicOr(flags, f);
}

struct LinkWs {
    double *W;           // [n][n] working distances (symmetric copy)
    int32_t *size;       // [n] cluster sizes (0 = dropped)
    int32_t *chain;      // [n]
    int32_t *next, *head, *tail;   // member lists: a merged cluster is the x list followed by the y list
    int32_t *m_head;     // [n-1] first leaf of the merge's x list
    int32_t *m_nx, *m_ny;
    double *m_h;         // [n-1] merge heights
    int32_t *leaf;       // [n] final list; every cluster that ever existed is a run of it
    int32_t *pos;        // [n] inverse of leaf
};

__global__ void __launch_bounds__(ORD_THREADS) linkage_chain_kernel(LinkWs ws, int n, const int32_t *flags) {
    __shared__ double s_v[32];
    __shared__ int s_k[32];
    const int tid = threadIdx.x;
    if (*flags) return;                              // an input scipy refuses: nothing is clustered
    const double INF = __longlong_as_double(0x7FF0000000000000ll);
    for (int i = tid; i < n; i += ORD_THREADS) {
        ws.size[i] = 1; ws.next[i] = -1; ws.head[i] = i; ws.tail[i] = i;
    }
    __syncthreads();
    int chain_len = 0, first_active = 0;
    for (int k = 0; k < n - 1; ++k) {
        if (chain_len == 0) {                        // the smallest live index starts a chain
            while (ws.size[first_active] == 0) ++first_active;      // uniform: every thread walks the same way
            if (tid == 0) ws.chain[0] = first_active;
            chain_len = 1;
            __syncthreads();
        }
        int x, y;
        double cur;
        while (true) {
            x = ws.chain[chain_len - 1];
            const double *row = ws.W + (size_t)x * n;
            double v = INF;
            int idx = 0x7FFFFFFF;
            for (int i = tid; i < n; i += ORD_THREADS) {
                if (i == x || ws.size[i] == 0) continue;
                const double dist = row[i];
                if (dist < v) { v = dist; idx = i; }         // ascending i inside a thread: first minimum kept
            }
            block_argmin(v, idx, s_v, s_k);
            if (chain_len > 1) {
                const int prev = ws.chain[chain_len - 2];
                const double pv = row[prev];
                if (!(v < pv)) { y = prev; cur = pv; break; }    // the previous element wins ties
            }
            if (tid == 0) ws.chain[chain_len] = idx;
            ++chain_len;
            __syncthreads();
        }
        chain_len -= 2;
        if (x > y) { const int t = x; x = y; y = t; }
        const int nx = ws.size[x], ny = ws.size[y];
        __syncthreads();
        if (tid == 0) {
            ws.m_head[k] = ws.head[x]; ws.m_nx[k] = nx; ws.m_ny[k] = ny; ws.m_h[k] = cur;
            ws.next[ws.tail[x]] = ws.head[y];
            ws.head[y] = ws.head[x];
            ws.size[x] = 0;
            ws.size[y] = nx + ny;
        }
        // Lance-Williams update for "average": (nx d_xi + ny d_yi) / (nx + ny), each operation rounded on its own
        const double fx = (double)nx, fy = (double)ny, fs = (double)(nx + ny);
        const double *rx = ws.W + (size_t)x * n;
        double *ry = ws.W + (size_t)y * n;
        for (int i = tid; i < n; i += ORD_THREADS) {
            if (i == x || i == y || ws.size[i] == 0) continue;
            const double nv = __ddiv_rn(__dadd_rn(__dmul_rn(fx, rx[i]), __dmul_rn(fy, ry[i])), fs);
            ry[i] = nv;
            ws.W[(size_t)i * n + y] = nv;
        }
        __syncthreads();
    }
    // the surviving list in order: leaf / pos (every cluster is a run of it)
    if (tid == 0) {
        int root = 0;
        while (ws.size[root] == 0) ++root;
        int node = ws.head[root];
        for (int p = 0; p < n; ++p) { ws.leaf[p] = node; ws.pos[node] = p; node = ws.next[node]; }
    }
}

// penalty[i][j] = height of the merge that joins i and j: merge k owns the rectangle (x run) x (y run).
__global__ void cophenetic_fill_kernel(LinkWs ws, int n, double *__restrict__ penalty, const int32_t *flags) {
    if (*flags) return;
    const int k = blockIdx.x;
    const int nx = ws.m_nx[k], ny = ws.m_ny[k];
    const int xs = ws.pos[ws.m_head[k]];
    const double h = ws.m_h[k];
    const long long total = (long long)nx * ny;
    for (long long t = (long long)blockIdx.y * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.y * blockDim.x) {
        const int a = ws.leaf[xs + (int)(t / ny)], b = ws.leaf[xs + nx + (int)(t % ny)];
        penalty[(size_t)a * n + b] = h;
        penalty[(size_t)b * n + a] = h;
    }
}

// normalizePenaltyMatrix's two sums (ordering/orderBySimilarity.py:50-53) run over the matrix row by row, one element
// after the other; a rounded fp64 sum depends on that order, so ONE thread adds while the block stages the next tile.
constexpr int SEQ_TILE = 2048;
__global__ void __launch_bounds__(256) sequential_sum_kernel(const double *__restrict__ a, const double *__restrict__ b,
                                                              long long total, double *__restrict__ sums) {
    __shared__ double buf[2][SEQ_TILE];
    const double *src = blockIdx.x == 0 ? a : b;
    const long long tiles = (total + SEQ_TILE - 1) / SEQ_TILE;
    double acc = 0.0;
    for (int t = threadIdx.x; t < SEQ_TILE; t += 256) buf[0][t] = t < total ? src[t] : 0.0;
    __syncthreads();
    for (long long tile = 0; tile < tiles; ++tile) {
        const int cur = (int)(tile & 1);
        const long long nb = (tile + 1) * SEQ_TILE;
        if (tile + 1 < tiles)
            for (int t = threadIdx.x; t < SEQ_TILE; t += 256) buf[cur ^ 1][t] = nb + t < total ? src[nb + t] : 0.0;
        if (threadIdx.x == 0) {
            const long long left = total - tile * SEQ_TILE;
            const int cnt = left < SEQ_TILE ? (int)left : SEQ_TILE;
            for (int t = 0; t < cnt; ++t) acc = __dadd_rn(acc, buf[cur][t]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[blockIdx.x] = acc;
}

// mode 0: out = d + strength * p  (search copy: strength 5, search/orderBySimilarity.py:40)
// mode 1: out = d + strength * (p * dsum / psum) when psum > 0, else d + strength * p  (ordering copy :55-62, :179-180)
// mode 2: out = p * dsum / psum when psum > 0, else p  (normalizePenaltyMatrix alone)
__global__ void penalized_distance_kernel(const double *__restrict__ d, const double *__restrict__ p, long long total,
                                          int mode, double strength, const double *__restrict__ sums,
                                          double *__restrict__ out) {
    double dsum = 0.0, psum = 0.0;
    if (mode >= 1) { dsum = sums[0]; psum = sums[1]; }
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        double pen = p[t];
        if (mode >= 1 && psum > 0) pen = __ddiv_rn(__dmul_rn(pen, dsum), psum);
        out[t] = mode == 2 ? pen : __dadd_rn(d[t], __dmul_rn(strength, pen));
    }
}

// reorderSymmetricMatrix: out[i][j] = out[j][i] = d[order[i]][order[j]] for j >= i (ordering copy, :237-245) or
// j > i with a zero diagonal (search copy, :105-113)
__global__ void reorder_symmetric_kernel(const double *__restrict__ d, int n, const int32_t *__restrict__ order,
                                         int zero_diagonal, double *__restrict__ out) {
    const long long total = (long long)n * n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / n), j = (int)(t % n);
        const int lo = i < j ? i : j, hi = i < j ? j : i;
        out[t] = (i == j && zero_diagonal) ? 0.0 : d[(size_t)order[lo] * n + order[hi]];
    }
}

// ------------------------------------------------------------------------------------------------------------
// One CTA per repetition: greedy insertion from the given starting order (mode bit 0), then 2-opt (mode bit 1).
struct PathWs {
    int32_t *path;       // [reps][2][n]
    double *edge;        // [reps][2][n]   edge[k] = d[path[k]][path[k+1]] of the current path
};

__global__ void __launch_bounds__(ORD_THREADS)
order_paths_kernel(const double *__restrict__ d, int n, const int32_t *__restrict__ start_orders, int mode, PathWs ws,
                   int32_t *__restrict__ orders, double *__restrict__ scores, double *__restrict__ reductions) {
    __shared__ double s_v[32];
    __shared__ int s_k[32];
    __shared__ double s_change;
    const int tid = threadIdx.x, rep = blockIdx.x;
    const double INF = __longlong_as_double(0x7FF0000000000000ll);
    const int32_t *start = start_orders + (size_t)rep * n;
    int32_t *pa = ws.path + (size_t)rep * 2 * n, *pb = pa + n;
    double *ea = ws.edge + (size_t)rep * 2 * n, *eb = ea + n;
    double score = 0.0;
    if (mode & 1) {
        if (tid == 0) {
            pa[0] = start[0]; pa[1] = start[1];
            ea[0] = d[(size_t)start[0] * n + start[1]];
            score = ea[0];
        }
        __syncthreads();
        for (int len = 2; len < n; ++len) {
            const int q = start[len];
            const double *rowq = d + (size_t)q * n;
            double bv = INF;
            int bp = 0x7FFFFFFF;
            if (tid == 0) {
                bv = rowq[pa[0]]; bp = 0;                               // at the beginning of the path
                const double ev = d[(size_t)pa[len - 1] * n + q];       // at the end: only if strictly better
                if (ev < bv) { bv = ev; bp = 1; }
            }
            for (int pos = 1 + tid; pos < len; pos += ORD_THREADS) {    // between path[pos - 1] and path[pos]
                const double v = __dadd_rn(__dadd_rn(d[(size_t)pa[pos - 1] * n + q], rowq[pa[pos]]), -ea[pos - 1]);
                if (v < bv) { bv = v; bp = pos + 1; }
            }
            block_argmin(bv, bp, s_v, s_k);
            const int ins = bp == 0 ? 0 : (bp == 1 ? len : bp - 1);
            for (int k = tid; k <= len; k += ORD_THREADS) {
                pb[k] = k < ins ? pa[k] : (k == ins ? q : pa[k - 1]);
                if (k < len) {
                    double e;
                    if (k < ins - 1) e = ea[k];
                    else if (k == ins - 1) e = d[(size_t)pa[ins - 1] * n + q];
                    else if (k == ins) e = rowq[pa[ins]];
                    else e = ea[k - 1];
                    eb[k] = e;
                }
            }
            if (tid == 0) score = __dadd_rn(score, bv);
            __syncthreads();
            int32_t *tp = pa; pa = pb; pb = tp;
            double *te = ea; ea = eb; eb = te;
        }
    } else {
        for (int k = tid; k < n; k += ORD_THREADS) pa[k] = start[k];
        __syncthreads();
    }
    double reduction = 0.0;
    if (mode & 2) {
        int32_t *o = pa;
        for (int i = 0; i + 2 < n; ++i) {
            int jmax = n - 1;
            while (jmax >= i + 2) {
                const int a = o[i], b = o[i + 1];
                const double *rowa = d + (size_t)a * n, *rowb = d + (size_t)b * n;
                const double dab = rowa[b];
                int found = -1;
                double change = 0.0;
                for (int top = jmax; top >= i + 2 && found < 0; top -= ORD_THREADS) {
                    const int j = top - tid;
                    int cand = -1;
                    if (j >= i + 2) {
                        const int oj1 = o[j - 1], oj = o[j];
                        change = __dadd_rn(__dadd_rn(__dadd_rn(rowa[oj1], rowb[oj]), -dab), -d[(size_t)oj1 * n + oj]);
                        if (change < -0.0000000000001) cand = j;
                    }
                    found = block_max(cand, s_k);
                    if (cand >= 0 && cand == found) s_change = change;
                }
                if (found < 0) break;
                __syncthreads();                                      // s_change written, every o[] read done
                if (tid == 0) reduction = __dadd_rn(reduction, s_change);
                const int seg = found - i - 1;                        // order[i+1 .. found-1] is reversed
                for (int t = tid; t < seg / 2; t += ORD_THREADS) {
                    const int u = o[i + 1 + t], w = o[found - 1 - t];
                    o[i + 1 + t] = w; o[found - 1 - t] = u;
                }
                __syncthreads();
                jmax = found - 1;
            }
        }
    }
    __syncthreads();
    for (int k = tid; k < n; k += ORD_THREADS) orders[(size_t)rep * n + k] = pa[k];
    if (tid == 0) { scores[rep] = score; reductions[rep] = reduction; }
}

}  // namespace fr3d
