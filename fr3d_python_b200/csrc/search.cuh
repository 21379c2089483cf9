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
// the reference's own.  One warp per occupied cube; hits are compacted with
// ballots and one atomicAdd per (nt1, segment).
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
    int32_t *cell_list;        // [n] table slots of the occupied cubes (any order)
    int32_t *n_cells;          // [1]
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
    if (t == 0) *ws.n_cells = 0;
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

// member-list allocation: a block-wide scan of the cell counts (and of the occupied-cube flags) and ONE atomicAdd per
// block on each of the two cursors (one per warp made this kernel atomic bound: ~28 k atomics on one address, 25 us)
constexpr int CELL_OFFSETS_THREADS = 1024;
__global__ void __launch_bounds__(CELL_OFFSETS_THREADS) cell_offsets_kernel(SearchWs ws, int64_t *counters) {
    __shared__ int s_cnt[32], s_occ[32];
    __shared__ long long s_base;
    __shared__ int s_cbase;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = (t <= ws.table_mask) ? ws.count[t] : 0;
    int incl = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xFFFFFFFFu, incl, off);
        if (lane >= off) incl += v;
    }
    const unsigned occ = __ballot_sync(0xFFFFFFFFu, c > 0);
    if (lane == 31) { s_cnt[warp] = incl; s_occ[warp] = __popc(occ); }
    __syncthreads();
    if (warp == 0) {                    // exclusive scan of the 32 warp totals
        const int wc = s_cnt[lane], wo = s_occ[lane];
        int ic = wc, io = wo;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xFFFFFFFFu, ic, off), u = __shfl_up_sync(0xFFFFFFFFu, io, off);
            if (lane >= off) { ic += v; io += u; }
        }
        s_cnt[lane] = ic - wc;
        s_occ[lane] = io - wo;
        if (lane == 31) {
            s_base = ic ? (long long)atomicAdd((unsigned long long *)&counters[3], (unsigned long long)ic) : 0;
            s_cbase = io ? atomicAdd(ws.n_cells, io) : 0;
        }
    }
    __syncthreads();
    if (c > 0) {
        ws.start[t] = (int32_t)(s_base + s_cnt[warp] + incl - c);
        // ... and the occupied cubes go on a compact list (the search runs one warp per cube)
        ws.cell_list[s_cbase + s_occ[warp] + __popc(occ & ((1u << lane) - 1u))] = (int32_t)t;
    }
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

// One warp per occupied cube.  The 14 cube lookups of the half stencil (hash probe, count, start) are done once per
// cube by 14 lanes, the member lists are concatenated into ONE candidate list, and the candidates' centre and
// identity fields are staged in shared memory once; every nt1 of the cube then scans that list in full chunks of
// 32 out of shared memory.  (One warp per nt1 repeated the probes and the gathers for each of the ~5 nts of a
// cube, and an atomicAdd per ballot on the single append cursor made it atomic bound: 655 us -> 321 us -> this.)
#ifndef FR3D_SEARCH_CAP
#define FR3D_SEARCH_CAP 64
#endif
#ifndef FR3D_SEARCH_MINB
#define FR3D_SEARCH_MINB 10      // A/B on the bench workload (cap, blocks): (128, 7) 254 us, (96, 9) 223, (64, 10) 214, (64, 12) 214
#endif
constexpr int SEARCH_CAP = FR3D_SEARCH_CAP;          // candidates staged per segment
constexpr int SEARCH_CHUNKS = SEARCH_CAP / 32;
constexpr int SEARCH_WARPS = 4;

struct SearchStage {
    double xyz[SEARCH_CAP][3];           // candidates (nt2)
    int32_t n2[SEARCH_CAP];
    int32_t chain[SEARCH_CAP], index[SEARCH_CAP], resid[SEARCH_CAP], alt[SEARCH_CAP];
    uint8_t s[SEARCH_CAP];
    double q_xyz[32][3];                 // a batch of up to 32 nt1 of the cube
    int32_t q_n1[32], q_chain[32], q_index[32], q_resid[32], q_alt[32];
    unsigned kept[32][SEARCH_CHUNKS];    // ballots of (nt1, chunk)
};

__global__ void __launch_bounds__(SEARCH_WARPS * 32, FR3D_SEARCH_MINB)
pair_search_kernel(fr3d_nts_t nts, SearchWs ws, double cutoff, int32_t *__restrict__ pair_i,
                   int32_t *__restrict__ pair_j, int32_t *__restrict__ pair_rank, int64_t pair_cap, int64_t *counters) {
    constexpr unsigned ALL = 0xFFFFFFFFu;
    __shared__ SearchStage stage[SEARCH_WARPS];
    SearchStage &S = stage[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int n_cells = *ws.n_cells;
    const int warps_total = gridDim.x * SEARCH_WARPS;
    for (int w = blockIdx.x * SEARCH_WARPS + (threadIdx.x >> 5); w < n_cells; w += warps_total) {
        const int slot1 = ws.cell_list[w];
        const int cnt1 = ws.count[slot1], beg1 = ws.start[slot1];
        const int first1 = ws.first[slot1];
        const int rep = ws.members[beg1];                   // any member: the cube's coordinates and group
        const int cx = ws.cell_xyz[rep * 3 + 0], cy = ws.cell_xyz[rep * 3 + 1], cz = ws.cell_xyz[rep * 3 + 2];
        const int group = nts.meta[(size_t)rep * FR3D_META_INTS + 2];
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

        for (int m0 = 0; m0 < cnt1; m0 += 32) {
            // a batch of nt1: lane l loads member m0 + l (n1 = -1: no glycosidic atom, NA:674)
            const int mm = min(32, cnt1 - m0);
            __syncwarp();
            if (lane < mm) {
                const int n1 = ws.members[beg1 + m0 + lane];
                const int32_t *m1 = nts.meta + (size_t)n1 * FR3D_META_INTS;
                S.q_n1[lane] = (nts.base_mask[n1] & 1u) ? n1 : -1;
                S.q_chain[lane] = m1[3]; S.q_index[lane] = m1[4]; S.q_resid[lane] = m1[5]; S.q_alt[lane] = m1[6];
                S.q_xyz[lane][0] = nts.center[(size_t)n1 * 3 + 0];
                S.q_xyz[lane][1] = nts.center[(size_t)n1 * 3 + 1];
                S.q_xyz[lane][2] = nts.center[(size_t)n1 * 3 + 2];
            }
            for (int sb = 0; sb < total; sb += SEARCH_CAP) {
                const int seg = min(SEARCH_CAP, total - sb);
                // stage the segment's candidates
                __syncwarp();
                for (int t0 = 0; t0 < seg; t0 += 32) {      // warp-uniform trip count: the shuffles need every lane
                    const int t = t0 + lane;
                    const int k = sb + t;
                    // stencil entry of candidate k = number of cubes whose (inclusive, non-decreasing) prefix end is <= k:
                    // a four-step bisection over the values held by lanes 0..12 (lanes 13.. hold the total)
                    int s = 0;
#pragma unroll
                    for (int step = 8; step >= 1; step >>= 1) {
                        const int e = __shfl_sync(ALL, end, min(s + step - 1, 13));
                        if (s + step <= 13 && k >= e) s += step;
                    }
                    const int sh = __shfl_sync(ALL, shift, s);
                    if (t < seg) {
                        const int n2 = ws.members[k + sh];
                        const int32_t *m2 = nts.meta + (size_t)n2 * FR3D_META_INTS;
                        S.s[t] = (uint8_t)s;
                        S.n2[t] = n2;
                        S.chain[t] = m2[3]; S.index[t] = m2[4]; S.resid[t] = m2[5]; S.alt[t] = m2[6];
                        S.xyz[t][0] = nts.center[(size_t)n2 * 3 + 0];
                        S.xyz[t][1] = nts.center[(size_t)n2 * 3 + 1];
                        S.xyz[t][2] = nts.center[(size_t)n2 * 3 + 2];
                    }
                }
                __syncwarp();
                // pass 1: every nt1 of the batch scans the staged candidates; ballots go to shared memory
                int nkept = 0;
                for (int m = 0; m < mm; ++m) {
                    const int n1 = S.q_n1[m];
                    const int chain1 = S.q_chain[m], index1 = S.q_index[m], resid1 = S.q_resid[m], alt1 = S.q_alt[m];
                    const double c1x = S.q_xyz[m][0], c1y = S.q_xyz[m][1], c1z = S.q_xyz[m][2];
#pragma unroll
                    for (int c = 0; c < SEARCH_CHUNKS; ++c) {
                        if (c * 32 >= seg) {                 // warp uniform: no candidates in this chunk
                            if (lane == 0) S.kept[m][c] = 0;
                            continue;
                        }
                        const int t = c * 32 + lane;
                        bool keep = false;
                        if (n1 >= 0 && t < seg) {
                            keep = true;
                            if (S.s[t] == 0) {              // NA:684-688
                                const int chain2 = S.chain[t];
                                if (chain1 > chain2) keep = false;
                                else if (chain1 == chain2 && index1 > S.index[t]) keep = false;
                            }
                            if (keep) {
                                const double dx = fabs(S.xyz[t][0] - c1x);
                                const double dy = fabs(S.xyz[t][1] - c1y);
                                const double dz = fabs(S.xyz[t][2] - c1z);
                                if (dx > cutoff || dy > cutoff || dz > cutoff) keep = false;         // NA:699-702
                                else if (resid1 == S.resid[t] && alt1 != S.alt[t]) keep = false;     // NA:706-713
                                else {
                                    // d = sqrt(d2) against the cutoff and 2.0 (NA:716-724): decided on d2 unless it is
                                    // within rounding of a threshold, where the reference's own sqrt decides
                                    const double d2 = dx * dx + dy * dy + dz * dz, c2 = cutoff * cutoff;
                                    if (d2 > c2 * (1.0 + 1e-12) || d2 < 4.0 * (1.0 - 1e-12)) keep = false;
                                    else if (!(d2 < c2 * (1.0 - 1e-12) && d2 > 4.0 * (1.0 + 1e-12))) {
                                        const double d = sqrt(d2);
                                        if (d > cutoff || d < 2.0) keep = false;
                                    }
                                }
                            }
                        }
                        const unsigned b = __ballot_sync(ALL, keep);
                        if (lane == 0) S.kept[m][c] = b;
                        nkept += __popc(b);
                    }
                }
                if (!nkept) continue;
                // one append for the whole (batch, segment), then pass 2 writes the pairs grouped by nt1
                long long basepos = 0;
                if (lane == 0) basepos = (long long)atomicAdd((unsigned long long *)&counters[0], (unsigned long long)nkept);
                basepos = __shfl_sync(ALL, basepos, 0);
                __syncwarp();
                for (int m = 0; m < mm; ++m) {
                    const int n1 = S.q_n1[m];
#pragma unroll
                    for (int c = 0; c < SEARCH_CHUNKS; ++c) {
                        const unsigned b = S.kept[m][c];
                        if ((b >> lane) & 1u) {
                            const int t = c * 32 + lane;
                            const long long pos = basepos + __popc(b & ((1u << lane) - 1u));
                            if (pos < pair_cap) {
                                pair_i[pos] = n1;
                                pair_j[pos] = S.n2[t];
                                pair_rank[pos] = first1 * 16 + (int)S.s[t];
                            }
                        }
                        basepos += __popc(b);
                    }
                }
            }
        }
    }
}

}  // namespace fr3d
