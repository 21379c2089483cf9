// FR3D search screening: all pairs of points of the same model closer than a cutoff.
//
// Replaces fixed_radius_search (fr3d/search/pair_processing.py:106-177), the geometric screen of the FR3D motif
// search (get_pairlist, :7-104, slices its distance-sorted output per pair of query positions).  Same method as
// the reference - buckets of side max_cutoff, the bucket itself plus 13 neighbours at the same z level or above -
// on the spatial hash of search.cuh: cells keyed (model, floor(x/d), floor(y/d), floor(z/d)) with Python's exact
// floor division, member lists by count -> allocate -> fill, one warp per (occupied cell, stencil entry).  Every
// pair carries the reference's ORIENTATION (pnt1 in the first bucket, pnt2 in the neighbour; pnt1 < pnt2 inside a
// bucket) and rank = first point of the first bucket * 16 + stencil entry, so sorting by (rank, pnt1, pnt2)
// reproduces the reference's list order.  The squared distance is the reference's expression evaluated without fma
// contraction: (x1-x2)^2 + (y1-y2)^2 + (z1-z2)^2, left to right.
#pragma once

#include "search.cuh"

namespace fr3d {

// x // d of Python floats: the floor of the EXACT quotient (floor(x / d) can be one too large when the rounded
// quotient lands on an integer)
__device__ __forceinline__ double py_floordiv(double x, double d) {
    double q = floor(x / d);
    const double r = fma(-q, d, x);              // exact remainder of the candidate
    if (r < 0.0) q -= 1.0;
    else if (r >= d) q += 1.0;
    return q;
}

__global__ void rs_assign_kernel(const double *__restrict__ xyz, const int32_t *__restrict__ model, int n_points,
                                 double d, SearchWs ws, int64_t *counters) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_points) return;
    int slot = -1;
    const double x = xyz[(size_t)n * 3], y = xyz[(size_t)n * 3 + 1], z = xyz[(size_t)n * 3 + 2];
    if (!(isnan(x) || isnan(y) || isnan(z))) {                      // pair_processing.py:137
        const double fx = py_floordiv(x, d), fy = py_floordiv(y, d), fz = py_floordiv(z, d);
        if (!(fabs(fx) < CELL_BIAS - 2 && fabs(fy) < CELL_BIAS - 2 && fabs(fz) < CELL_BIAS - 2)) {
            atomicOr((unsigned long long *)&counters[2], 1ull);     // FR3D_E_RANGE, reported by the host
        } else {
            const int cx = (int)fx, cy = (int)fy, cz = (int)fz;
            ws.cell_xyz[n * 3 + 0] = cx; ws.cell_xyz[n * 3 + 1] = cy; ws.cell_xyz[n * 3 + 2] = cz;
            const unsigned long long key = cell_key(model[n], cx, cy, cz);
            uint32_t h = hash64(key) & ws.table_mask;
            while (true) {
                const unsigned long long prev = atomicCAS(&ws.keys[h], HASH_EMPTY, key);
                if (prev == HASH_EMPTY || prev == key) break;
                h = (h + 1) & ws.table_mask;
            }
            slot = (int)h;
            atomicAdd(&ws.count[h], 1);
            atomicMin(&ws.first[h], n);
        }
    }
    ws.cell_slot[n] = slot;
}

__global__ void rs_fill_kernel(int n_points, SearchWs ws) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_points) return;
    const int slot = ws.cell_slot[n];
    if (slot >= 0) ws.members[ws.start[slot] + atomicAdd(&ws.cursor[slot], 1)] = n;
}

// pair_processing.py:147-151: neighbours at the same z level or above; entry 13 = the bucket itself
__constant__ int8_t c_rs_stencil[14][3] = {{1, 1, 1}, {0, 1, 1}, {-1, 1, 1}, {1, 0, 1}, {0, 0, 1}, {-1, 0, 1}, {1, -1, 1},
                                           {0, -1, 1}, {-1, -1, 1}, {1, 1, 0}, {0, 1, 0}, {-1, 1, 0}, {1, 0, 0}, {0, 0, 0}};

constexpr int RS_WARPS = 4;

__global__ void __launch_bounds__(RS_WARPS * 32)
rs_pairs_kernel(const double *__restrict__ xyz, const int32_t *__restrict__ model, SearchWs ws, double d_squared,
                int32_t *__restrict__ pair_a, int32_t *__restrict__ pair_b, int32_t *__restrict__ pair_rank,
                double *__restrict__ dist2, int64_t pair_cap, int64_t *counters) {
    constexpr unsigned ALL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const long long n_tasks = (long long)(*ws.n_cells) * 14;
    const long long warps_total = (long long)gridDim.x * RS_WARPS;
    for (long long task = (long long)blockIdx.x * RS_WARPS + (threadIdx.x >> 5); task < n_tasks; task += warps_total) {
        const int j = (int)(task % 14);
        const int slot1 = ws.cell_list[task / 14];
        const int cnt1 = ws.count[slot1], beg1 = ws.start[slot1], first1 = ws.first[slot1];
        int slot2 = slot1;
        if (j < 13) {
            const int rep = ws.members[beg1];
            slot2 = hash_find(ws, cell_key(model[rep], ws.cell_xyz[rep * 3 + 0] + c_rs_stencil[j][0],
                                           ws.cell_xyz[rep * 3 + 1] + c_rs_stencil[j][1],
                                           ws.cell_xyz[rep * 3 + 2] + c_rs_stencil[j][2]));
            if (slot2 < 0) continue;
        }
        const int cnt2 = ws.count[slot2], beg2 = ws.start[slot2];
        for (int m = 0; m < cnt1; ++m) {
            const int p1 = ws.members[beg1 + m];
            const double x1 = xyz[(size_t)p1 * 3], y1 = xyz[(size_t)p1 * 3 + 1], z1 = xyz[(size_t)p1 * 3 + 2];
            for (int t0 = 0; t0 < cnt2; t0 += 32) {
                const int t = t0 + lane;
                bool keep = false;
                int p2 = 0;
                double dd = 0.0;
                if (t < cnt2) {
                    p2 = ws.members[beg2 + t];
                    const double dx = x1 - xyz[(size_t)p2 * 3], dy = y1 - xyz[(size_t)p2 * 3 + 1], dz = z1 - xyz[(size_t)p2 * 3 + 2];
                    dd = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                    keep = dd < d_squared && (j < 13 || p1 < p2);                 // :163, :171-174
                }
                const unsigned b = __ballot_sync(ALL, keep);
                if (!b) continue;
                long long base = 0;
                if (lane == 0) base = (long long)atomicAdd((unsigned long long *)&counters[0], (unsigned long long)__popc(b));
                base = __shfl_sync(ALL, base, 0);
                const long long pos = base + __popc(b & ((1u << lane) - 1u));
                if (keep && pos < pair_cap) {
                    pair_a[pos] = p1; pair_b[pos] = p2;
                    pair_rank[pos] = first1 * 16 + j;
                    dist2[pos] = dd;
                }
            }
        }
    }
}

}  // namespace fr3d
