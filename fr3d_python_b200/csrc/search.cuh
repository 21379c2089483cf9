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
    int4 *cell_info;           // [T] (cx, cy, cz, group) of the cube in the slot (written by the nt that claimed it)
    double2 *sorted;           // [n][3] the search's view of an nt, in member order (cube by cube): (x, y), (z, n | glycosidic
                               //        atom present), (chain, index, residue id, alt id) - 48 contiguous bytes per candidate
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
                if (prev == HASH_EMPTY) ws.cell_info[h] = make_int4(cx, cy, cz, group);
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
        // the record the pair search reads: the members of a cube are neighbours in memory, so a warp's 32 candidates are
        // a handful of contiguous runs instead of 32 scattered centre / meta rows (it was L2-sector bound on those gathers)
        const int4 *m = reinterpret_cast<const int4 *>(nts.meta + (size_t)n * FR3D_META_INTS);
        const int4 lo = m[0], hi = m[1];                    // meta[3..6] = chain, index, residue id, alt id
        double2 *rec = ws.sorted + (size_t)pos * 3;
        rec[0] = make_double2(nts.center[(size_t)n * 3 + 0], nts.center[(size_t)n * 3 + 1]);
        rec[1] = make_double2(nts.center[(size_t)n * 3 + 2],
                              __hiloint2double((nts.base_mask[n] & 1u) ? 1 : 0, n));
        const int4 key = make_int4(lo.w, hi.x, hi.y, hi.z);
        rec[2] = *reinterpret_cast<const double2 *>(&key);
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
// cube by 14 lanes and the member lists are treated as ONE candidate list.  The nt1 of the cube (a batch of up to 32)
// wait in shared memory; the candidates are taken 32 at a time, one per lane, and stay in REGISTERS while the lane
// tests its candidate against every nt1 of the batch (four broadcast shared-memory loads per nt1) and collects the
// accepting nt1 as a bit mask.  One atomicAdd per (batch, 32 candidates) places the pairs, written grouped by nt1.
// History on the bench workload: one warp per nt1 655 us -> one append per warp 321 us -> one warp per cube with the
// candidates staged in shared memory and a ballot per (nt1, 32 candidates) 166 us -> this (see profiles/README.md).
#ifndef FR3D_SEARCH_MINB
#define FR3D_SEARCH_MINB 8
#endif
constexpr int SEARCH_WARPS = 4;
#ifndef FR3D_SEARCH_GROUP
#define FR3D_SEARCH_GROUP 1      // 2 / 4 x 32 candidates per atomicAdd were slower (0.183 / 0.188 vs 0.177 ms, profiles/ab_search_r2.log)
#endif
constexpr int SEARCH_GROUP = FR3D_SEARCH_GROUP;      // x 32 candidates per append

struct SearchBatch {                     // the nt1 of a cube, per warp
    double xy[32][2];
    double z[32];
    int4 key[32];                        // chain, index, residue id, alt id
    int32_t n1[32];                      // -1: no glycosidic atom (NA:674)
};

__global__ void __launch_bounds__(SEARCH_WARPS * 32, FR3D_SEARCH_MINB)
pair_search_kernel(fr3d_nts_t nts, SearchWs ws, double cutoff, int32_t *__restrict__ pair_i,
                   int32_t *__restrict__ pair_j, int32_t *__restrict__ pair_rank, int64_t pair_cap, int64_t *counters) {
    constexpr unsigned ALL = 0xFFFFFFFFu;
    __shared__ __align__(16) SearchBatch batch[SEARCH_WARPS];
    SearchBatch &Q = batch[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int n_cells = *ws.n_cells;
    const int warps_total = gridDim.x * SEARCH_WARPS;
    const double c2 = cutoff * cutoff;
    const double c2_hi = c2 * (1.0 + 1e-12), c2_lo = c2 * (1.0 - 1e-12), four_hi = 4.0 * (1.0 + 1e-12), four_lo = 4.0 * (1.0 - 1e-12);
    for (int w = blockIdx.x * SEARCH_WARPS + (threadIdx.x >> 5); w < n_cells; w += warps_total) {
        const int slot1 = ws.cell_list[w];
        const int cnt1 = ws.count[slot1], beg1 = ws.start[slot1];
        const int first1 = ws.first[slot1];
        const int4 info = ws.cell_info[slot1];              // the cube's coordinates and group: one load, no walk
        const int cx = info.x, cy = info.y, cz = info.z, group = info.w;
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
            // a batch of nt1: lane l loads member m0 + l
            const int mm = min(32, cnt1 - m0);
            __syncwarp();
            if (lane < mm) {
                const double2 *rec = ws.sorted + (size_t)(beg1 + m0 + lane) * 3;
                const double2 xy = rec[0], zn = rec[1], kk = rec[2];
                Q.n1[lane] = __double2hiint(zn.y) ? __double2loint(zn.y) : -1;
                Q.key[lane] = *reinterpret_cast<const int4 *>(&kk);
                Q.xy[lane][0] = xy.x;
                Q.xy[lane][1] = xy.y;
                Q.z[lane] = zn.x;
            }
            __syncwarp();
            // the candidates in groups of SEARCH_GROUP x 32 (a cube's whole stencil fits one group on RNA): ONE atomicAdd per
            // group - ncu showed the kernel waiting on same-address atomics with one per 32 candidates (45 % of the stall
            // samples, ~0.5 atomics per clock on one L2 address)
            for (int sb = 0; sb < total; sb += 32 * SEARCH_GROUP) {
                unsigned accept[SEARCH_GROUP];              // bit m: nt1 m of the batch pairs with this lane's candidate c
                int n2[SEARCH_GROUP], st[SEARCH_GROUP];
                int mine = 0;
#pragma unroll
                for (int c = 0; c < SEARCH_GROUP; ++c) {
                    accept[c] = 0; n2[c] = 0; st[c] = 0;
                    if (sb + 32 * c >= total) continue;     // warp uniform
                    const int k = sb + 32 * c + lane;
                    const bool valid = k < total;
                    // stencil entry of candidate k = number of cubes whose (inclusive, non-decreasing) prefix end is <= k:
                    // a four-step bisection over the values held by lanes 0..12 (lanes 13.. hold the total)
                    int s = 0;
#pragma unroll
                    for (int step = 8; step >= 1; step >>= 1) {
                        const int e = __shfl_sync(ALL, end, min(s + step - 1, 13));
                        if (s + step <= 13 && k >= e) s += step;
                    }
                    const int sh = __shfl_sync(ALL, shift, s);
                    int4 k2 = make_int4(0, 0, 0, 0);
                    double x2 = 0.0, y2 = 0.0, z2 = 0.0;
                    if (valid) {
                        const double2 *rec = ws.sorted + (size_t)(k + sh) * 3;
                        const double2 xy = rec[0], zn = rec[1], kk = rec[2];
                        n2[c] = __double2loint(zn.y);
                        k2 = *reinterpret_cast<const int4 *>(&kk);
                        x2 = xy.x; y2 = xy.y; z2 = zn.x;
                    }
                    st[c] = s;
                    // The test of one (nt1, candidate): branch-free, decided on the squared distance; a pair whose d2 is
                    // within rounding of a threshold goes on the `band` mask and is settled below by the reference's own
                    // sqrt comparison (NA:716-724).  The per-axis screen of NA:699-702 is implied: |dx| > cutoff gives
                    // d >= |dx| > cutoff (sqrt(x * x) == |x| in IEEE arithmetic), so the distance test rejects the pair too.
                    unsigned acc = 0, band = 0;
                    for (int m = 0; m < mm; ++m) {
                        const int n1 = Q.n1[m];
                        const int4 k1 = Q.key[m];
                        const double2 xy1 = *reinterpret_cast<const double2 *>(Q.xy[m]);
                        const double z1 = Q.z[m];
                        const double dx = x2 - xy1.x, dy = y2 - xy1.y, dz = z2 - z1;
                        const double d2 = dx * dx + dy * dy + dz * dz;
                        const bool order_ok = (s != 0) | (k1.x < k2.x) | ((k1.x == k2.x) & (k1.y <= k2.y));   // NA:684-688
                        const bool alt_ok = (k1.z != k2.z) | (k1.w == k2.w);                                   // NA:706-713
                        const bool maybe = valid & (n1 >= 0) & order_ok & alt_ok & (d2 <= c2_hi) & (d2 >= four_lo);
                        const bool sure = (d2 < c2_lo) & (d2 > four_hi);
                        acc |= ((maybe & sure) ? 1u : 0u) << m;
                        band |= ((maybe & !sure) ? 1u : 0u) << m;
                    }
                    while (band) {                          // rare: within 1e-12 (relative) of 2 A or of the cutoff
                        const int m = __ffs(band) - 1;
                        band &= band - 1;
                        const double dx = x2 - Q.xy[m][0], dy = y2 - Q.xy[m][1], dz = z2 - Q.z[m];
                        const double d = sqrt(dx * dx + dy * dy + dz * dz);
                        if (!(d > cutoff || d < 2.0)) acc |= 1u << m;
                    }
                    accept[c] = acc;
                    mine += __popc(acc);
                }
                // the pairs are written grouped by nt1 (the screen kernel of K4 reads the nt1 record from global memory: runs
                // of one nt1 stay in L1 - candidate-major order made that kernel 3 % slower, profiles/ab_search_r2.log)
                const int nkept = __reduce_add_sync(ALL, mine);
                if (!nkept) continue;
                long long pos = 0;
                if (lane == 0) pos = (long long)atomicAdd((unsigned long long *)&counters[0], (unsigned long long)nkept);
                pos = __shfl_sync(ALL, pos, 0);
                for (int m = 0; m < mm; ++m) {
                    const int n1 = Q.n1[m];
#pragma unroll
                    for (int c = 0; c < SEARCH_GROUP; ++c) {
                        if (sb + 32 * c >= total) continue;
                        const unsigned bm = __ballot_sync(ALL, (accept[c] >> m) & 1u);
                        if ((bm >> lane) & 1u) {
                            const long long at = pos + __popc(bm & ((1u << lane) - 1u));
                            if (at < pair_cap) {
                                pair_i[at] = n1;
                                pair_j[at] = n2[c];
                                pair_rank[at] = first1 * 16 + st[c];
                            }
                        }
                        pos += __popc(bm);
                    }
                }
            }
        }
    }
}

}  // namespace fr3d
