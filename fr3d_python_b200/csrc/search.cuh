// K2 + K3: candidate-pair search on a uniform spatial hash of base centres.
//
// Replaces make_nt_cubes_half (NA_pairwise_interactions.py:410-443) and the
// loop / screens of annotate_nt_nt_interactions (NA:661-724).
//
// Cells are 12 A cubes keyed (group, floor(x/12), floor(y/12), floor(z/12)),
// group = (structure, model).  The hash table is open addressing in global
// memory (it lives in L2 for every configuration we run); cell member lists
// are built with count -> allocate -> fill, and the search walks the
// reference's 14-cell half stencil so that the ORIENTATION of every pair and
// the visit-order rank (first nt seen in the cell of nt1, stencil index) are
// the reference's own.  One warp per nt1; hits are compacted with ballots
// and one atomicAdd per warp.
#pragma once

#include "tables.cuh"

namespace fr3d {

static constexpr unsigned long long HASH_EMPTY = 0xFFFFFFFFFFFFFFFFull;
static constexpr int CELL_BIAS = 2048;   // cells in [-2048, 2047] on each axis (+-24.5 k Angstrom)

struct SearchWs {
    unsigned long long *keys;  // [T]
    int32_t *count;            // [T]
    int32_t *first;            // [T] smallest nt index in the cell = first-seen order of the cube
    int32_t *start;            // [T]
    int32_t *cursor;           // [T]
    int32_t *members;          // [n]
    int32_t *cell_slot;        // [n] (-1: nt not in any cube)
    int32_t *cell_xyz;         // [n][3]
    uint32_t table_mask;       // T - 1
};

__device__ __forceinline__ unsigned long long cell_key(int group, int cx, int cy, int cz) {
    return ((unsigned long long)(uint32_t)group << 36) | ((unsigned long long)(uint32_t)(cx + CELL_BIAS) << 24) |
           ((unsigned long long)(uint32_t)(cy + CELL_BIAS) << 12) | (unsigned long long)(uint32_t)(cz + CELL_BIAS);
}

__device__ __forceinline__ uint32_t hash64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (uint32_t)k;
}

__global__ void search_reset_kernel(SearchWs ws) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t <= ws.table_mask) {
        ws.keys[t] = HASH_EMPTY;
        ws.count[t] = 0;
        ws.first[t] = 0x7FFFFFFF;
        ws.start[t] = 0;
        ws.cursor[t] = 0;
    }
}

// make_nt_cubes_half, NA:424-436: floor(coord / 12) in fp64, insert into the cube.
__global__ void cell_assign_kernel(fr3d_nts_t nts, SearchWs ws, double cell_size, int64_t *counters) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nts.n_nt) return;
    int slot = -1;
    if (nts.base_mask[n] & FR3D_MASK_HAS_FRAME) {
        const double *c = nts.center + (size_t)n * 3;
        const double fx = floor(c[0] / cell_size), fy = floor(c[1] / cell_size), fz = floor(c[2] / cell_size);
        if (!(fabs(fx) < CELL_BIAS - 2 && fabs(fy) < CELL_BIAS - 2 && fabs(fz) < CELL_BIAS - 2)) {
            atomicOr((unsigned long long *)&counters[2], 1ull);   // FR3D_E_RANGE, reported by the host
        } else {
            const int cx = (int)fx, cy = (int)fy, cz = (int)fz;
            ws.cell_xyz[n * 3 + 0] = cx; ws.cell_xyz[n * 3 + 1] = cy; ws.cell_xyz[n * 3 + 2] = cz;
            const int group = nts.meta[n * FR3D_META_INTS + 2];
            const unsigned long long key = cell_key(group, cx, cy, cz);
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

// member-list allocation: a warp scan of the cell counts and ONE atomicAdd per warp on the cursor
__global__ void cell_offsets_kernel(SearchWs ws, int64_t *counters) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int c = (t <= ws.table_mask) ? ws.count[t] : 0;
    int incl = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xFFFFFFFFu, incl, off);
        if (lane >= off) incl += v;
    }
    const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    if (!total) return;
    long long base = 0;
    if (lane == 31) base = (long long)atomicAdd((unsigned long long *)&counters[3], (unsigned long long)total);
    base = __shfl_sync(0xFFFFFFFFu, base, 31);
    if (c > 0) ws.start[t] = (int32_t)(base + incl - c);
}

__global__ void cell_fill_kernel(fr3d_nts_t nts, SearchWs ws) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nts.n_nt) return;
    const int slot = ws.cell_slot[n];
    if (slot >= 0) {
        const int pos = ws.start[slot] + atomicAdd(&ws.cursor[slot], 1);
        ws.members[pos] = n;
    }
}

__device__ __forceinline__ int hash_find(const SearchWs &ws, unsigned long long key) {
    uint32_t h = hash64(key) & ws.table_mask;
    while (true) {
        const unsigned long long k = ws.keys[h];
        if (k == key) return (int)h;
        if (k == HASH_EMPTY) return -1;
        h = (h + 1) & ws.table_mask;
    }
}

// NA:438 — same cube and 13 neighbours, no two in opposite directions
__constant__ int8_t c_half_stencil[14][3] = {{0, 0, 0}, {0, 0, 1}, {0, 1, 0}, {0, 1, 1}, {0, 1, -1}, {1, 0, 0}, {1, 0, 1},
                                             {1, 0, -1}, {1, 1, 0}, {1, 1, 1}, {1, 1, -1}, {1, -1, 0}, {1, -1, 1}, {1, -1, -1}};

// One warp per nt1.  The 14 cube lookups (hash probe, count, start) are done by 14 lanes at once, the member
// lists of the cubes are then walked as ONE concatenated candidate list in full chunks of 32, and the kept pairs
// of up to SEARCH_CHUNKS chunks are appended with a single atomicAdd (the append cursor is one address: an atomic
// per ballot made it the bottleneck of this kernel).
constexpr int SEARCH_CHUNKS = 8;

__global__ void __launch_bounds__(128)
pair_search_kernel(fr3d_nts_t nts, SearchWs ws, double cutoff, int32_t *__restrict__ pair_i,
                   int32_t *__restrict__ pair_j, int32_t *__restrict__ pair_rank, int64_t pair_cap, int64_t *counters) {
    constexpr unsigned ALL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int n1 = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (n1 >= nts.n_nt) return;
    const int slot1 = ws.cell_slot[n1];
    if (slot1 < 0) return;                              // no base centre (NA:666)
    if (!(nts.base_mask[n1] & 1u)) return;              // nt1 lacks its glycosidic atom (NA:674)
    const int32_t *m1 = nts.meta + (size_t)n1 * FR3D_META_INTS;
    const int group = m1[2], chain1 = m1[3], index1 = m1[4], resid1 = m1[5], alt1 = m1[6];
    const double c1x = nts.center[n1 * 3 + 0], c1y = nts.center[n1 * 3 + 1], c1z = nts.center[n1 * 3 + 2];
    const int cx = ws.cell_xyz[n1 * 3 + 0], cy = ws.cell_xyz[n1 * 3 + 1], cz = ws.cell_xyz[n1 * 3 + 2];
    const int first1 = ws.first[slot1];

    // lane s < 14: the cube of stencil entry s (a cube that was never made has no members, NA:663)
    int cnt = 0, beg = 0;
    if (lane < 14) {
        const int slot2 = lane == 0 ? slot1
                                    : hash_find(ws, cell_key(group, cx + c_half_stencil[lane][0],
                                                             cy + c_half_stencil[lane][1], cz + c_half_stencil[lane][2]));
        if (slot2 >= 0) { cnt = ws.count[slot2]; beg = ws.start[slot2]; }
    }
    int end = cnt;                                      // inclusive prefix: candidates [end - cnt, end) are cube s's
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) {
        const int v = __shfl_up_sync(ALL, end, off);
        if (lane >= off) end += v;
    }
    const int total = __shfl_sync(ALL, end, 13);
    const int shift = beg - (end - cnt);                // member index = candidate index + shift

    for (int sb = 0; sb < total; sb += 32 * SEARCH_CHUNKS) {
        int n2v[SEARCH_CHUNKS];
        unsigned kept[SEARCH_CHUNKS];
        unsigned spack = 0;                             // 4 bits per chunk: stencil entry of this lane's candidate
        int nkept = 0;
#pragma unroll
        for (int c = 0; c < SEARCH_CHUNKS; ++c) {
            kept[c] = 0;
            n2v[c] = -1;
            if (sb + c * 32 >= total) continue;         // warp uniform
            const int k = sb + c * 32 + lane;
            bool keep = false;
            int s = 0;
#pragma unroll
            for (int j = 0; j < 13; ++j) s += (k >= __shfl_sync(ALL, end, j)) ? 1 : 0;
            const int sh = __shfl_sync(ALL, shift, s < 14 ? s : 13);
            if (k < total) {
                const int n2 = ws.members[k + sh];
                n2v[c] = n2;
                const int32_t *m2 = nts.meta + (size_t)n2 * FR3D_META_INTS;
                keep = true;
                if (s == 0) {                           // NA:684-688
                    const int chain2 = m2[3];
                    if (chain1 > chain2) keep = false;
                    else if (chain1 == chain2 && index1 > m2[4]) keep = false;
                }
                if (keep) {
                    const double dx = fabs(nts.center[n2 * 3 + 0] - c1x);
                    const double dy = fabs(nts.center[n2 * 3 + 1] - c1y);
                    const double dz = fabs(nts.center[n2 * 3 + 2] - c1z);
                    if (dx > cutoff || dy > cutoff || dz > cutoff) keep = false;             // NA:699-702
                    else if (resid1 == m2[5] && alt1 != m2[6]) keep = false;                 // NA:706-713
                    else {
                        const double d = sqrt(dx * dx + dy * dy + dz * dz);                  // NA:716
                        if (d > cutoff || d < 2.0) keep = false;                             // NA:719-724
                    }
                }
            }
            spack |= (unsigned)s << (4 * c);
            kept[c] = __ballot_sync(ALL, keep);
            nkept += __popc(kept[c]);
        }
        if (!nkept) continue;
        long long basepos = 0;
        if (lane == 0) basepos = (long long)atomicAdd((unsigned long long *)&counters[0], (unsigned long long)nkept);
        basepos = __shfl_sync(ALL, basepos, 0);
#pragma unroll
        for (int c = 0; c < SEARCH_CHUNKS; ++c) {
            if ((kept[c] >> lane) & 1u) {
                const long long pos = basepos + __popc(kept[c] & ((1u << lane) - 1u));
                if (pos < pair_cap) {
                    pair_i[pos] = n1;
                    pair_j[pos] = n2v[c];
                    pair_rank[pos] = first1 * 16 + (int)((spack >> (4 * c)) & 15u);
                }
            }
            basepos += __popc(kept[c]);
        }
    }
}

}  // namespace fr3d
