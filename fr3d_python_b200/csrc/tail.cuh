// A8 on the device: the order-dependent tail of annotate_nt_nt_interactions for every structure of a batch.
//
// Replaces (fr3d/classifiers/NA_pairwise_interactions.py)
//   the visit-order recording of per-pair results                 NA:758-1013
//   check_for_two_interactions_on_same_edge                       NA:528-632
//   the emission of the surviving basepairs                       NA:1026-1044
//   calculate_crossing_numbers                                    NA:1057-1179
//   annotate_covalent_connections                                 NA:1181-1210
//
// Hit records arrive unordered (K4 appends them with atomics).  They are binned by structure with the sort key
// (rank, nt1, nt2, slot) that reproduces the reference's cube walk - in ONE pass into segments whose capacity is the
// structure's share of hit_cap by nt count (tail_bin_kernel, one atomic per warp and structure); only when a structure
// has more hits than its share do the exact scan -> scatter kernels run (they return at once otherwise);
// then ONE thread block per structure does everything else on integer ids:
//   sort the structure's hits                       -> the order in which the reference records results
//   walk them: non-basepair hits append (label, a, b) rows; basepair hits add one entry to each partner's list
//   sort the entries by (first appearance of the nt, sequence) = the iteration order of unit_id_to_basepairs
//   same-edge cleanup (sequential by construction, but only over nts with two or more basepairs), emission
//   nested-cWW selection per chain id (greedy over pairs sorted by span; one warp per chain), crossing numbers
//   rank the labels by first appearance, sort the rows by (label rank, sequence), expand the reversed duplicates
//   covalent links: sort the nts by (model symmetry chain, index), link successive keys
// and writes the final rows (label id, nt1, nt2, crossing) in the reference's append order.
// Sorting: a bitonic network with the keys in registers (<= TAIL_SORT_CAP keys, 256-thread blocks) or a block-wide LSD radix
// sort (everything larger, and the 1024-thread blocks of the large structures); all keys are unique, so the result is
// deterministic.
#pragma once

#include "tables.cuh"

namespace fr3d {

#ifndef FR3D_TAIL_MINB
#define FR3D_TAIL_MINB 4             // blocks per SM the structure kernel is compiled for (4: 64 registers, no spills)
#endif
constexpr int TAIL_THREADS = 256;        // block size for ordinary structures (four blocks per SM)
constexpr int TAIL_THREADS_MANY = 128;   // ... when a call has more structures than 256-thread blocks fit at once (eight per SM)
constexpr int TAIL_THREADS_BIG = 1024;   // ... for structures of TAIL_BIG_NT nts and more: one block per SM, four times the lanes
constexpr int TAIL_BIG_NT = 1024;
constexpr int TAIL_SORT_CAP = 2048;      // elements sorted in registers by 256 threads (exchange buffer: 16 KB of shared memory)
constexpr int TAIL_MAX_LABELS = 1024;
constexpr int TAIL_MAX_NT_BITS = 17;     // structures of up to 131 072 nts
constexpr int TAIL_INDEX_BITS = 23;
constexpr uint32_t TAIL_NONE = 0xFFFFFFFFu;
constexpr int TAIL_F_REMOVED = 1, TAIL_F_NEAR = 2;
template <int TT>
__host__ __device__ constexpr size_t tail_smem_bytes() { return (size_t)8 * TT * sizeof(unsigned long long); }
constexpr size_t TAIL_SMEM = tail_smem_bytes<TAIL_THREADS>();

struct TailWs {
    int32_t *seg_cnt;             // [S]     hits per structure
    int32_t *seg_off;             // [S + 1]
    int32_t *seg_cur;             // [S]     = the structure's hit count after tail_bin_kernel
    int32_t *nt_struct;           // [n]     structure of every nt
    unsigned long long *tmpk;     // [2H]    radix sort: the other buffer of the ping-pong (keys)
    uint32_t *tmpv;               // [2H]    ... (values)
    unsigned long long *tmpc;     // [n]     ... for the covalent keys (one per nt)
    int32_t *overflow;            // [1]     a structure's hits did not fit its share: exact scan + scatter take over
    unsigned long long *hkey;     // [H]     hits of a structure: sort key; later the nested-cWW candidates
    uint32_t *hval;               // [H]     hit index
    unsigned long long *ekey;     // [2H]    basepair entries: (first appearance of u, sequence)
    uint32_t *eval;               // [2H]    hit index * 2 + (1: the entry recorded for u2)
    int32_t *e_u, *e_v, *e_lab;   // [2H]    entries in iteration order
    uint8_t *e_flag;              // [2H]
    unsigned long long *akey;     // [2H]    appended rows: sort keys (label rank, append index)
    uint32_t *aval;               // [2H]    scratch: entries in recording order
    int32_t *alab, *aa, *ab;      // [2H]
    int32_t *across;              // [2H]
    uint32_t *first;              // [n]     first appearance of the nt in unit_id_to_basepairs
    int32_t *gstart, *gcount;     // [n]     its entries
    unsigned long long *ckey;     // [n]     covalent sort
    int32_t *ends;                // [n_ends] nested_cWW_endpoints of every chain
};

typedef unsigned long long tail_u64;

// -DFR3D_TAIL_PROFILE: thread 0 of every block adds the cycles of each phase to tc[8 + phase] (tools/profile_tail.py)
#ifdef FR3D_TAIL_PROFILE
#define TAIL_PHASE(k)                                                                 \
    do {                                                                              \
        __syncthreads();                                                              \
        if (threadIdx.x == 0) {                                                       \
            const long long now_ = clock64();                                         \
            atomicAdd((tail_u64 *)&tc[8 + (k)], (tail_u64)(now_ - phase_t0));         \
            phase_t0 = now_;                                                          \
        }                                                                             \
    } while (0)
#else
#define TAIL_PHASE(k)
#endif

__device__ __forceinline__ int tail_find_struct(const int32_t *__restrict__ off, int n_struct, int nt) {
    int lo = 0, hi = n_struct;                    // largest s with off[s] <= nt
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= nt) lo = mid; else hi = mid;
    }
    return lo;
}

// per structure: zero cursors, segment = the structure's share of hit_cap by nt count; per nt: its structure
__global__ void tail_reset_kernel(TailWs ws, fr3d_tail_ident_t id, int n_nt, int64_t hit_cap, int64_t *tc) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_struct = id.n_struct;
    if (t <= n_struct) {
        if (t < n_struct) { ws.seg_cnt[t] = 0; ws.seg_cur[t] = 0; }
        const long long total = id.struct_off[n_struct];
        ws.seg_off[t] = total > 0 ? (int32_t)((hit_cap * (long long)id.struct_off[t]) / total) : 0;
    }
    if (t < n_nt) ws.nt_struct[t] = n_struct > 0 ? tail_find_struct(id.struct_off, n_struct, t) : 0;
    if (t == 0) { tc[0] = 0; tc[1] = 0; tc[2] = 0; tc[3] = 0; *ws.overflow = 0; }
}

__device__ __forceinline__ tail_u64 tail_hit_key(const fr3d_hit_t &hit, int32_t index_offset, int lo) {
    // (rank, nt1, nt2, slot): rank = first nt of nt1's cube * 16 + stencil entry (search.cuh), all local to the structure
    const int i = hit.i - index_offset, j = hit.j - index_offset;
    const tail_u64 rank = (tail_u64)(uint32_t)(hit.rank - 16 * index_offset - 16 * lo);
    return (rank << (2 * TAIL_MAX_NT_BITS + 4)) | ((tail_u64)(uint32_t)(i - lo) << (TAIL_MAX_NT_BITS + 4)) |
           ((tail_u64)(uint32_t)(j - lo) << 4) | (tail_u64)hit.slot;
}

// One pass: hit -> its structure's segment.  Hits of one cube sit next to each other in the stream, so the lanes of a
// warp mostly agree on the structure: one atomicAdd per (warp, structure).  seg_cur[s] ends as the exact hit count of
// the structure whether or not its hits fitted.
__global__ void tail_bin_kernel(const fr3d_hit_t *__restrict__ hits, int64_t hit_cap, const int64_t *counters,
                                int32_t index_offset, fr3d_tail_ident_t id, TailWs ws, int64_t *tc) {
    const long long n_hits = min((long long)counters[1], (long long)hit_cap);
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long h0 = (long long)blockIdx.x * blockDim.x + threadIdx.x - lane; h0 < n_hits; h0 += stride) {
        const long long h = h0 + lane;
        const bool live = h < n_hits;
        fr3d_hit_t hit;
        int s = -1 - lane;                                   // idle lanes: a group of their own
        if (live) {
            hit = hits[h];
            s = ws.nt_struct[hit.i - index_offset];
        }
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, s);
        if (!live) continue;
        const int leader = __ffs(peers) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&ws.seg_cur[s], __popc(peers));
        base = __shfl_sync(peers, base, leader);
        const int at = base + __popc(peers & ((1u << lane) - 1u));
        const int lo = id.struct_off[s];
        if (id.struct_off[s + 1] - lo > (1 << TAIL_MAX_NT_BITS)) { atomicOr((tail_u64 *)&tc[2], 1ull); continue; }
        const int seg = ws.seg_off[s];
        if (at >= ws.seg_off[s + 1] - seg) { *ws.overflow = 1; continue; }
        ws.hkey[seg + at] = tail_hit_key(hit, index_offset, lo);
        ws.hval[seg + at] = (uint32_t)h;
    }
}

// exclusive scan of the per-structure hit counts (one block; the structure count is small next to the hit count)
__global__ void __launch_bounds__(1024) tail_scan_kernel(TailWs ws, int n_struct) {
    __shared__ int s_w[32];
    __shared__ int s_run;
    if (!*ws.overflow) return;                   // every structure fitted its share: the segments stand as they are
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int base = 0; base < n_struct; base += 1024) {
        const int t = base + threadIdx.x;
        const int c = t < n_struct ? ws.seg_cur[t] : 0;      // exact counts left by tail_bin_kernel
        int incl = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xFFFFFFFFu, incl, off);
            if (lane >= off) incl += v;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = s_w[lane];
            int iw = w;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int v = __shfl_up_sync(0xFFFFFFFFu, iw, off);
                if (lane >= off) iw += v;
            }
            s_w[lane] = iw - w;
        }
        __syncthreads();
        const int run = s_run;
        if (t < n_struct) ws.seg_off[t] = run + s_w[warp] + incl - c;
        __syncthreads();
        if (threadIdx.x == 1023) s_run = run + s_w[warp] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) ws.seg_off[n_struct] = s_run;
}

__global__ void tail_scatter_kernel(const fr3d_hit_t *__restrict__ hits, int64_t hit_cap, const int64_t *counters,
                                    int32_t index_offset, fr3d_tail_ident_t id, TailWs ws, int64_t *tc) {
    if (!*ws.overflow) return;
    const long long n_hits = min((long long)counters[1], (long long)hit_cap);
    for (long long h = (long long)blockIdx.x * blockDim.x + threadIdx.x; h < n_hits; h += (long long)gridDim.x * blockDim.x) {
        const fr3d_hit_t hit = hits[h];
        const int s = ws.nt_struct[hit.i - index_offset];
        const int lo = id.struct_off[s];
        if (id.struct_off[s + 1] - lo > (1 << TAIL_MAX_NT_BITS)) continue;        // flagged by tail_bin_kernel
        const int pos = ws.seg_off[s] + atomicAdd(&ws.seg_cnt[s], 1);
        ws.hkey[pos] = tail_hit_key(hit, index_offset, lo);
        ws.hval[pos] = (uint32_t)h;
    }
}

// ---- block-wide primitives -----------------------------------------------------------------------------------------
// exclusive scan of one value per thread; every thread gets the block total too
template <int TT>
__device__ __forceinline__ int tail_block_scan(int v, int *s_w, int &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int u = __shfl_up_sync(0xFFFFFFFFu, incl, off);
        if (lane >= off) incl += u;
    }
    __syncthreads();                         // s_w may still be read from the previous call
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    int before = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < TT / 32; ++w) {
        const int x = s_w[w];
        if (w < warp) before += x;
        tot += x;
    }
    total = tot;
    return before + incl - v;
}

// Bitonic sort of up to 256 * E keys held in registers, E per thread, element i = tid * E + e; the payload rides in the
// low bits of the key (32- or 64-bit keys).  Compare distance j < E: inside the thread; E <= j < 32 E: with a lane of
// the own warp (shuffles); larger: through shared memory.  All loops have compile-time bounds, so the E elements never
// leave the registers.  Missing elements (i >= n) are +infinity.  (The first version kept (key, value) pairs in shared
// memory with one comparator per thread and step: ~120 instructions per comparator step and thread, 87 k cycles per
// 1 k-element sort with seven blocks per SM; registers + shuffles: 47 k; key-only with narrow keys: see profiles/.)
template <int TT, int E, typename K>
__device__ __forceinline__ void tail_sort_keys_regs(K *keys, int n, K *s_buf) {
    constexpr int P = TT * E;
    constexpr unsigned ALL = 0xFFFFFFFFu;
    const int tid = threadIdx.x;
    K k[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = tid * E + e;
        k[e] = i < n ? keys[i] : (K)~(K)0;
    }
#pragma unroll
    for (int size = 2; size <= P; size <<= 1) {
#pragma unroll
        for (int j = size >> 1; j >= 1; j >>= 1) {
            if (j >= 32 * E) {
                __syncthreads();
#pragma unroll
                for (int e = 0; e < E; ++e) s_buf[tid * E + e] = k[e];
                __syncthreads();
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int i = tid * E + e;
                    const K pk = s_buf[i ^ j];
                    const bool take_min = ((i & j) == 0) == ((i & size) == 0);
                    if (take_min ? pk < k[e] : pk > k[e]) k[e] = pk;
                }
            } else if (j >= E) {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int i = tid * E + e;
                    const K pk = __shfl_xor_sync(ALL, k[e], j / E);
                    const bool take_min = ((i & j) == 0) == ((i & size) == 0);
                    if (take_min ? pk < k[e] : pk > k[e]) k[e] = pk;
                }
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    if ((e & j) == 0) {
                        const int i = tid * E + e;
                        const bool asc = (i & size) == 0;
                        if ((k[e] > k[e | j]) == asc) { const K t = k[e]; k[e] = k[e | j]; k[e | j] = t; }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = tid * E + e;
        if (i < n) keys[i] = k[e];
    }
}

// ---- block-wide LSD radix sort (8-bit digits, stable) ------------------------------------------------------------------
// The sorts of the tail have short keys (22 to ~57 significant bits: label rank | sequence, rank | nt1 | nt2 | slot | position
// ...), so three to eight counting passes beat the O(n log^2 n) comparator network even at n ~ 800, and by an order of
// magnitude for the large structures.  Stability comes from giving every WARP a contiguous segment of the input: per-warp
// digit counts (s_hist[warp][digit], no atomics: a warp's leaders of distinct digits write distinct words), one exclusive
// scan in (digit, warp) order, then each warp scatters its segment in order, 32 keys at a time, the lanes of equal digit
// ranked by __match_any_sync.  Keys ping-pong between `keys` and `tmp` in global memory (L2-resident); the number of passes
// comes from the OR of all keys.  s_hist: (TT / 32) * 256 ints of the dynamic shared memory (8 KB / 32 KB).
// the lanes holding the same 8-bit digit as this one (live lanes only): eight ballots - __match_any_sync takes one hardware
// step per DISTINCT value, up to 32 with random digits
__device__ __forceinline__ unsigned tail_digit_peers(int d, bool live) {
    unsigned peers = __ballot_sync(0xFFFFFFFFu, live);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const unsigned vote = __ballot_sync(0xFFFFFFFFu, (d >> b) & 1);
        peers &= ((d >> b) & 1) ? vote : ~vote;
    }
    return peers;
}

template <int TT, typename K, bool KV>
__device__ __noinline__ void tail_radix_sort(K *keys, uint32_t *vals, K *tmpk, uint32_t *tmpv, int n, void *smem,
                                             int skip_low = 0) {      // skip_low: payload bits that need no ordering
    constexpr unsigned ALL = 0xFFFFFFFFu;
    constexpr int W = TT / 32;
    __shared__ unsigned s_or[2];
    __shared__ int s_scan[W];
    int *s_hist = reinterpret_cast<int *>(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __syncthreads();
    if (n <= 1) return;
    if (tid < 2) s_or[tid] = 0;
    __syncthreads();
    {
        unsigned lo = 0, hi = 0;
        for (int i = tid; i < n; i += TT) { const unsigned long long k = (unsigned long long)keys[i]; lo |= (unsigned)k; hi |= (unsigned)(k >> 32); }
        lo = __reduce_or_sync(ALL, lo); hi = __reduce_or_sync(ALL, hi);
        if (lane == 0) { if (lo) atomicOr(&s_or[0], lo); if (hi) atomicOr(&s_or[1], hi); }
    }
    __syncthreads();
    const int bits = s_or[1] ? 64 - __clz(s_or[1]) : (s_or[0] ? 32 - __clz(s_or[0]) : 0);
    const int passes = bits > skip_low ? (bits - skip_low + 7) >> 3 : 0;
    const int seg = (n + W - 1) / W;                       // warp w owns [w * seg, (w + 1) * seg)
    const int w_lo = warp * seg, w_hi = min(n, w_lo + seg);
    K *src = keys, *dst = tmpk;
    uint32_t *vsrc = vals, *vdst = tmpv;
    for (int pass = 0; pass < passes; ++pass) {
        const int shift = skip_low + pass * 8;
        for (int e = tid; e < W * 256; e += TT) s_hist[e] = 0;
        __syncthreads();
        // (the key of the next step is loaded one step ahead: the steps of a warp are serial, the loads need not be)
        K knext = w_lo + lane < w_hi ? src[w_lo + lane] : (K)0;
        for (int i0 = w_lo; i0 < w_hi; i0 += 32) {
            const int i = i0 + lane;
            const K k = knext;
            if (i + 32 < w_hi) knext = src[i + 32];
            const bool live = i < w_hi;
            const int d = (int)((k >> shift) & 255);
            const unsigned peers = tail_digit_peers(d, live);
            if (live && lane == __ffs(peers) - 1) s_hist[warp * 256 + d] += __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        // exclusive scan in (digit, warp) order: entry e = digit * W + warp; thread t owns the 8 entries 8t .. 8t + 7
        int run[8], tot = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int e = tid * 8 + q;
            run[q] = tot;
            tot += s_hist[(e % W) * 256 + e / W];
        }
        int block_total;
        const int before = tail_block_scan<TT>(tot, s_scan, block_total);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int e = tid * 8 + q;
            s_hist[(e % W) * 256 + e / W] = before + run[q];
        }
        __syncthreads();
        knext = w_lo + lane < w_hi ? src[w_lo + lane] : (K)0;
        uint32_t vnext = 0;
        if (KV && w_lo + lane < w_hi) vnext = vsrc[w_lo + lane];
        for (int i0 = w_lo; i0 < w_hi; i0 += 32) {
            const int i = i0 + lane;
            const bool live = i < w_hi;
            const K k = knext;
            const uint32_t v = vnext;
            if (i + 32 < w_hi) { knext = src[i + 32]; if (KV) vnext = vsrc[i + 32]; }
            const int d = (int)((k >> shift) & 255);
            const unsigned peers = tail_digit_peers(d, live);
            if (live) {
                const int at = s_hist[warp * 256 + d] + __popc(peers & ((1u << lane) - 1u));
                dst[at] = k;
                if (KV) vdst[at] = v;
            }
            __syncwarp();
            if (live && lane == __ffs(peers) - 1) s_hist[warp * 256 + d] += __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        { K *t = src; src = dst; dst = t; }
        { uint32_t *t = vsrc; vsrc = vdst; vdst = t; }
    }
    if (src != keys) {
        for (int i = tid; i < n; i += TT) { keys[i] = src[i]; if (KV) vals[i] = vsrc[i]; }
    }
    __syncthreads();
}

#ifndef FR3D_TAIL_RADIX_ALWAYS
#define FR3D_TAIL_RADIX_ALWAYS 0     // A/B: 1 = no register network at all
#endif
// Up to 2 048 keys in a 256-thread block: the register network (no memory traffic at all; the radix sort's passes cost
// more there: sort hits 43 k -> 53 k cycles per 316-nt structure, profiles/tail_phases_r2e_c4_radix_everywhere.txt).
// Everything larger, and everything in the 1 024-thread blocks of the large structures: the radix sort (C3: 64 k hits in
// 0.3 M cycles instead of 17.7 M through the global-memory network, 3.4 M with register tiles).
template <int TT, typename K>
__device__ __forceinline__ void tail_block_sort_keys(K *keys, int n, void *smem, void *tmp, int skip_low = 0) {
    if (TT != TAIL_THREADS_BIG && n <= 8 * TT && !FR3D_TAIL_RADIX_ALWAYS) {
        __syncthreads();
        if (n <= 1) return;
        K *s_buf = reinterpret_cast<K *>(smem);
        if (n <= TT) tail_sort_keys_regs<TT, 1, K>(keys, n, s_buf);
        else if (n <= 2 * TT) tail_sort_keys_regs<TT, 2, K>(keys, n, s_buf);
        else if (n <= 4 * TT) tail_sort_keys_regs<TT, 4, K>(keys, n, s_buf);
        else tail_sort_keys_regs<TT, 8, K>(keys, n, s_buf);
        __syncthreads();
    } else {
        tail_radix_sort<TT, K, false>(keys, nullptr, reinterpret_cast<K *>(tmp), nullptr, n, smem, skip_low);
    }
}

struct TailEntrySrc {      // what one basepair hit says about the entry recorded for u1 (d = 0) or u2 (d = 1)
    int u, v, lab;
};

__device__ __forceinline__ TailEntrySrc tail_entry_of(const fr3d_hit_t &h, int d, int base, const fr3d_tail_tables_t &tb) {
    const int a = h.i - base, b = h.j - base;
    const int near = h.code & 1;
    const int u1 = (h.code & 2) ? b : a, u2 = (h.code & 2) ? a : b;      // orientation 21 was used: nt2 is u1 (NA:981)
    TailEntrySrc e;
    e.u = d ? u2 : u1;
    e.v = d ? u1 : u2;
    e.lab = tb.boxes[h.box].label[d][near];
    return e;
}

// One block per structure (dynamic assignment through tc[1] / tc[3]).  BIG = false: blocks of 256 threads, four per SM, take
// the structures below TAIL_BIG_NT nts (all of them when the caller does not announce large ones) - or blocks of 128
// threads, eight per SM, when the call has more structures than 4 x SMs: a structure then takes longer, but all of them
// are in flight at once instead of in 1.7 rounds (1000 structures: 0.35 -> 0.30 ms; a pipeline chunk of ~100 structures is
// faster with the larger blocks, profiles/ab_tail_threads_r2.log); BIG = true: blocks of
// 1024 threads, one per SM, take the others - the phases are block-wide networks and scans whose time falls with the lane
// count, and a large structure has nobody to share the SM with anyway.
template <int TT, int MINB, bool BIG>
__global__ void __launch_bounds__(TT, MINB)
tail_structure_kernel(fr3d_nts_t nts, fr3d_tail_ident_t id, fr3d_tail_tables_t tb, const fr3d_hit_t *__restrict__ hits,
                      int32_t index_offset, TailWs ws, fr3d_row_t *__restrict__ rows, int64_t row_cap,
                      int32_t *__restrict__ row_start, int32_t *__restrict__ row_count, int64_t *tc) {
    extern __shared__ __align__(16) unsigned char tail_smem[];
    __shared__ uint32_t s_lfirst[TAIL_MAX_LABELS];
    __shared__ uint16_t s_lrank[TAIL_MAX_LABELS];
    __shared__ int s_w[TT / 32];
    __shared__ int s_struct, s_nent, s_napp, s_nnest, s_nmulti, s_err;
    __shared__ long long s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int T = TT;
    constexpr size_t SMEM_BYTES = tail_smem_bytes<TT>();
    const int n_labels = tb.n_labels;

    while (true) {
        __syncthreads();
        if (tid == 0) {
            s_struct = (int)atomicAdd((tail_u64 *)&tc[BIG ? 3 : 1], 1ull);
            s_nent = 0; s_napp = 0; s_nnest = 0; s_nmulti = 0; s_err = 0;
        }
        __syncthreads();
        const int s = s_struct;
        if (s >= id.n_struct) break;
#ifdef FR3D_TAIL_PROFILE
        long long phase_t0 = clock64();
#endif
        const int lo = id.struct_off[s], n = id.struct_off[s + 1] - lo;
        if (id.max_struct_nt >= TAIL_BIG_NT && (n >= TAIL_BIG_NT) != BIG) continue;      // the other kernel's structure
        const int hb = ws.seg_off[s], H = ws.seg_cur[s];
        const int base = index_offset + lo;                                  // hit.i - base = position in the structure
        if (n > (1 << TAIL_MAX_NT_BITS)) {                                   // flagged by the scatter kernel
            if (tid == 0) { row_start[s] = -1; row_count[s] = 0; }
            continue;
        }
        tail_u64 *hkey = ws.hkey + hb;
        uint32_t *hval = ws.hval + hb;
        const size_t eb = 2 * (size_t)hb;
        tail_u64 *ekey = ws.ekey + eb, *akey = ws.akey + eb;
        uint32_t *eval = ws.eval + eb, *aval = ws.aval + eb;
        int32_t *e_u = ws.e_u + eb, *e_v = ws.e_v + eb, *e_lab = ws.e_lab + eb;
        uint8_t *e_flag = ws.e_flag + eb;
        int32_t *alab = ws.alab + eb, *aa = ws.aa + eb, *ab = ws.ab + eb, *across = ws.across + eb;
        uint32_t *first = ws.first + lo;
        int32_t *gstart = ws.gstart + lo, *gcount = ws.gcount + lo;
        const int32_t *meta = nts.meta + (size_t)lo * FR3D_META_INTS;
        const int32_t *nt_chain = id.nt_chain + lo;

        // ---- init, and the ranges the keys below rely on
        const int c_lo = id.struct_chain_off[s], c_hi = id.struct_chain_off[s + 1];
        if (c_hi - c_lo > 65536 || n_labels > TAIL_MAX_LABELS) s_err = 1;
        for (int u = tid; u < n; u += T) {
            first[u] = TAIL_NONE; gcount[u] = 0;
            const int ix = meta[(size_t)u * FR3D_META_INTS + 4], cov = id.nt_cov[lo + u], ch = nt_chain[u];
            if (ix < 0 || ix >= (1 << TAIL_INDEX_BITS) || cov < 0 || cov >= (1 << 23) || ch < c_lo || ch >= c_hi) s_err = 1;
            else if (ix >= id.chain_ends_off[ch + 1] - id.chain_ends_off[ch]) s_err = 1;
        }
        for (int l = tid; l < min(n_labels, TAIL_MAX_LABELS); l += T) s_lfirst[l] = TAIL_NONE;
        __syncthreads();
        if (s_err) {
            if (tid == 0) { row_start[s] = -1; row_count[s] = 0; atomicOr((tail_u64 *)&tc[2], 2ull); }
            continue;
        }
        // nested_cWW_endpoints of the structure's chains (NA:1106): in shared memory (the sort buffer, free between the
        // nested-cWW sort and the final sort) when they fit, else in the global workspace
        const int ends_lo = id.chain_ends_off[c_lo];
        const int ends_len = id.chain_ends_off[c_hi] - ends_lo;
        const bool ends_shared = ends_len <= (int)(SMEM_BYTES / sizeof(int32_t));
        int32_t *ends_base = ends_shared ? reinterpret_cast<int32_t *>(tail_smem) : ws.ends + ends_lo;

        TAIL_PHASE(0);
        // ---- the reference's visit order.  When the bits allow (always, for structures whose hits fit the register
        // sort) the key is re-packed as (rank, nt1, nt2, slot, position in the segment) with field widths from n and H,
        // and sorted WITHOUT a value array: half the shuffles and selects of a (key, value) sort.
        int nb = 1;
        while ((1 << nb) < n) ++nb;
        int hbits = 1;
        while ((1 << hbits) < H) ++hbits;
        const bool packed = 3 * nb + 8 + hbits <= 64;
        const tail_u64 hmask = (1ull << hbits) - 1ull;
        if (packed) {
            const tail_u64 fm = (1ull << TAIL_MAX_NT_BITS) - 1ull;
            for (int t = tid; t < H; t += T) {
                const tail_u64 k = hkey[t];              // (rank : nt1 : nt2 : slot) in fixed 17-bit fields (tail_scatter_kernel)
                const tail_u64 slot = k & 15ull, j = (k >> 4) & fm, i = (k >> (TAIL_MAX_NT_BITS + 4)) & fm,
                               rank = k >> (2 * TAIL_MAX_NT_BITS + 4);
                hkey[t] = ((((rank << nb | i) << nb | j) << 4 | slot) << hbits) | (tail_u64)t;
            }
            tail_block_sort_keys<TT, tail_u64>(hkey, H, tail_smem, ws.tmpk + eb, hbits);      // (the position is payload)
        } else {
            tail_radix_sort<TT, tail_u64, true>(hkey, hval, ws.tmpk + eb, ws.tmpv + eb, H, tail_smem);
        }
        // hit index of sorted position p
#define TAIL_HIT_OF(p) (packed ? hval[(int)(hkey[p] & hmask)] : hval[p])

        TAIL_PHASE(1);
        // ---- walk: positions by block scans, so that the appended rows and the basepair entries come out in the
        // reference's recording order (their index IS their sequence number)
        uint32_t *ent_hit = aval;                        // entry e (recording order) -> hit index * 2 + (1: recorded for u2)
        {
            int run_app = 0, run_ent = 0;
            for (int p0 = 0; p0 < H; p0 += T) {
                const int p = p0 + tid;
                int cnt_app = 0, cnt_ent = 0, l0 = -1, l1 = -1, a = 0, b = 0;
                uint32_t hidx = 0;
                bool swap = false;
                TailEntrySrc e0 = {0, 0, 0};
                if (p < H) {
                    hidx = TAIL_HIT_OF(p);
                    const fr3d_hit_t h = hits[hidx];
                    a = h.i - base; b = h.j - base;
                    if (h.slot == FR3D_SLOT_BP) {
                        e0 = tail_entry_of(h, 0, base, tb);
                        cnt_ent = 2;
                    } else if (h.slot < FR3D_SLOT_BP) {
                        const int32_t *sl = tb.slot_label + ((int)h.slot * 32 + (h.code & 31)) * 2;
                        l0 = sl[0]; l1 = sl[1];
                        swap = h.slot == FR3D_SLOT_SO21 || h.slot == FR3D_SLOT_SR21;       // recorded for (nt2, nt1)
                        if (l0 < 0 || l0 >= n_labels || l1 >= n_labels) { s_err = 1; l0 = -1; }
                        else cnt_app = l1 >= 0 ? 2 : 1;
                    }
                }
                int total;
                const int ex = tail_block_scan<TT>(cnt_app | (cnt_ent << 16), s_w, total);
                if (cnt_app) {
                    const int k = run_app + (ex & 0xFFFF);
                    alab[k] = l0; aa[k] = swap ? b : a; ab[k] = swap ? a : b;
                    atomicMin(&s_lfirst[l0], (uint32_t)k);
                    if (cnt_app == 2) {                                                       // NA:765 / NA:775
                        alab[k + 1] = l1; aa[k + 1] = swap ? a : b; ab[k + 1] = swap ? b : a;
                        atomicMin(&s_lfirst[l1], (uint32_t)(k + 1));
                    }
                }
                if (cnt_ent) {
                    const int e = run_ent + (ex >> 16);
                    ent_hit[e] = hidx * 2u;
                    ent_hit[e + 1] = hidx * 2u + 1u;
                    ekey[e] = (tail_u64)(uint32_t)e0.u;                   // the nt for now; its first appearance below
                    ekey[e + 1] = (tail_u64)(uint32_t)e0.v;
                    atomicMin(&first[e0.u], (uint32_t)e);
                    atomicMin(&first[e0.v], (uint32_t)(e + 1));
                    atomicAdd(&gcount[e0.u], 1);
                    atomicAdd(&gcount[e0.v], 1);
                }
                run_app += total & 0xFFFF;
                run_ent += total >> 16;
            }
            if (tid == 0) { s_napp = run_app; s_nent = run_ent; }
        }
        __syncthreads();
        const int nent = s_nent;
        const int napp_walk = s_napp;

        TAIL_PHASE(2);
        // ---- unit_id_to_basepairs in iteration order: nts by first appearance, entries by sequence (NA:1001-1013);
        // key = (entry index of the nt's first entry, entry index), sorted without values
        for (int e = tid; e < nent; e += T) ekey[e] = ((tail_u64)first[(int)ekey[e]] << 32) | (tail_u64)(uint32_t)e;
        tail_block_sort_keys<TT, tail_u64>(ekey, nent, tail_smem, ws.tmpk + eb);
        for (int q = tid; q < nent; q += T) {
            const uint32_t ev = ent_hit[(int)(ekey[q] & 0xFFFFFFFFull)];
            eval[q] = ev;
            const TailEntrySrc e = tail_entry_of(hits[ev >> 1], (int)(ev & 1u), base, tb);
            e_u[q] = e.u; e_v[q] = e.v; e_lab[q] = e.lab; e_flag[q] = 0;
            if (e.lab < 0 || e.lab >= n_labels) s_err = 1;
        }
        __syncthreads();
        // group starts, and the list (in iteration order, in the low half of akey) of the nts whose basepairs can
        // conflict at all: two entries on the same edge sharing hydrogen-bond atoms.  That test ignores the marks made
        // so far, which can only suppress comparisons, so an nt that fails it never marks anything - and it runs in
        // parallel, one thread per nt, while the order-dependent pass below only visits the nts that pass.
        int32_t *multi = reinterpret_cast<int32_t *>(akey);
        {
            int running = 0;
            for (int q0 = 0; q0 < nent; q0 += T) {
                const int q = q0 + tid;
                int flag = 0;
                if (q < nent) {
                    const int u = e_u[q];
                    if (q == 0 || e_u[q - 1] != u) {
                        gstart[u] = q;
                        const int len = gcount[u];
                        for (int i = q; i < q + len - 1 && !flag; ++i) {
                            const uint32_t f1 = tb.labels[e_lab[i]].flags;
                            const fr3d_hit_t *h1 = hits + (eval[i] >> 1);
                            const tail_u64 at1 = (eval[i] & 1u) ? tb.boxes[h1->box].atoms2 : tb.boxes[h1->box].atoms1;
                            for (int j = i + 1; j < q + len; ++j) {
                                const uint32_t f2 = tb.labels[e_lab[j]].flags;
                                if (((f1 >> 8) & 0xFF) != ((f2 >> 8) & 0xFF)) continue;
                                const fr3d_hit_t *h2 = hits + (eval[j] >> 1);
                                const tail_u64 common = at1 & ((eval[j] & 1u) ? tb.boxes[h2->box].atoms2 : tb.boxes[h2->box].atoms1);
                                if (common && ((common & tb.hydrogen_atoms) || __popcll(common) > 1)) { flag = 1; break; }
                            }
                        }
                    }
                }
                int total;
                const int ex = tail_block_scan<TT>(flag, s_w, total);
                if (flag) multi[running + ex] = q;
                running += total;
            }
            if (tid == 0) s_nmulti = running;
        }
        __syncthreads();

        TAIL_PHASE(3);
        // ---- check_for_two_interactions_on_same_edge (NA:528-632): order dependent, one thread, conflicts are rare
        if (tid == 0 && !s_err) {
            const int nmulti = s_nmulti;
            for (int m = 0; m < nmulti; ++m) {
                const int g0 = multi[m];
                const int u = e_u[g0];
                const int len = gcount[u];
                for (int i = g0; i < g0 + len - 1; ++i) {
                    if (e_flag[i] & TAIL_F_REMOVED) continue;                              // NA:545
                    const int lab1 = e_lab[i], v1 = e_v[i];
                    const uint32_t f1 = tb.labels[lab1].flags;
                    const fr3d_hit_t *h1 = hits + (eval[i] >> 1);
                    const tail_u64 at1 = (eval[i] & 1u) ? tb.boxes[h1->box].atoms2 : tb.boxes[h1->box].atoms1;
                    const double cd1 = h1->cutoff_distance, gap1 = h1->max_gap;
                    for (int j = i + 1; j < g0 + len; ++j) {
                        if (e_flag[j] & (TAIL_F_REMOVED | TAIL_F_NEAR)) continue;          // NA:551-557
                        const int lab2 = e_lab[j], v2 = e_v[j];
                        const uint32_t f2 = tb.labels[lab2].flags;
                        if (((f1 >> 8) & 0xFF) != ((f2 >> 8) & 0xFF)) continue;            // different edges
                        const fr3d_hit_t *h2 = hits + (eval[j] >> 1);
                        const tail_u64 at2 = (eval[j] & 1u) ? tb.boxes[h2->box].atoms2 : tb.boxes[h2->box].atoms1;
                        const tail_u64 common = at1 & at2;
                        if (!common) continue;
                        if (!((common & tb.hydrogen_atoms) || __popcll(common) > 1)) continue;
                        const double cd2 = h2->cutoff_distance, gap2 = h2->max_gap;
                        const bool n1 = f1 & 2u, n2 = f2 & 2u;
                        int flag = 0, target = -1;                                          // partner whose pair is marked
                        if (n1 && n2) {
                            if (cd1 < cd2) { if (cd2 > 0.5) { flag = TAIL_F_REMOVED; target = v2; } }
                            else if (cd1 > 0.5) { flag = TAIL_F_REMOVED; target = v1; }
                        } else if (!n1 && !n2) {
                            flag = TAIL_F_NEAR;
                            target = gap1 < gap2 ? v2 : v1;
                        } else {
                            if (cd2 > 0.5) { flag = TAIL_F_REMOVED; target = v2; }
                            else if (cd1 > 0.5) { flag = TAIL_F_REMOVED; target = v1; }
                        }
                        if (flag) {         // the sets hold (u, target) and (target, u): every entry with either key
                            for (int q = g0; q < g0 + len; ++q)
                                if (e_v[q] == target) e_flag[q] |= (uint8_t)flag;
                            const int t0 = gstart[target], tl = gcount[target];
                            for (int q = t0; q < t0 + tl; ++q)
                                if (e_v[q] == u) e_flag[q] |= (uint8_t)flag;
                        }
                    }
                }
            }
        }
        __syncthreads();

        TAIL_PHASE(4);
        // ---- emission (NA:1026-1044): the first entry of an unordered pair in iteration order, unless removed; the
        // emitted rows follow the walk's rows in entry order (block scan)
        {
            int running = napp_walk;
            for (int q0 = 0; q0 < nent; q0 += T) {
                const int q = q0 + tid;
                int lab = -1, u = 0, v = 0;
                if (q < nent) {
                    const int fl = e_flag[q];
                    if (!(fl & TAIL_F_REMOVED)) {
                        u = e_u[q]; v = e_v[q];
                        bool is_first = true;
                        for (int q2 = gstart[u]; q2 < q; ++q2)
                            if (e_v[q2] == v) is_first = false;
                        const int t0 = gstart[v], t1 = t0 + gcount[v];
                        for (int q2 = t0; q2 < t1 && q2 < q; ++q2)
                            if (e_v[q2] == u) is_first = false;
                        if (is_first) {
                            lab = e_lab[q];
                            if (lab >= 0 && lab < n_labels && (fl & TAIL_F_NEAR)) {
                                lab = tb.labels[lab].near;
                                if (lab < 0 || lab >= n_labels) s_err = 1;
                            }
                            if (lab < 0 || lab >= n_labels) lab = -1;
                        }
                    }
                }
                int total;
                const int ex = tail_block_scan<TT>(lab >= 0 ? 1 : 0, s_w, total);
                if (lab >= 0) {
                    const int k = running + ex;
                    alab[k] = lab; aa[k] = u; ab[k] = v;
                    atomicMin(&s_lfirst[lab], (uint32_t)k);
                }
                running += total;
            }
            if (tid == 0) s_napp = running;
        }
        __syncthreads();
        const int napp = s_napp;

        TAIL_PHASE(5);
        // ---- nested cWW pairs (NA:1077-1131): candidates keyed (chain, span, start), greedy in that order per chain
        tail_u64 *nkey = hkey;           // the hit keys are no longer needed
        for (int k = tid; k < napp; k += T) {
            if (!(tb.labels[alab[k]].flags & 4u)) continue;
            const int a = aa[k], b = ab[k];
            if (nt_chain[a] != nt_chain[b]) continue;
            const int pa = meta[(size_t)a * FR3D_META_INTS], pb = meta[(size_t)b * FR3D_META_INTS];
            // AU UA CG GC GU UG (NA:1089); parents A C G U = 0 1 2 3
            const bool ok = (pa == 0 && pb == 3) || (pa == 3 && pb == 0) || (pa == 1 && pb == 2) || (pa == 2 && pb == 1) ||
                            (pa == 2 && pb == 3) || (pa == 3 && pb == 2);
            if (!ok) continue;
            const int ia = meta[(size_t)a * FR3D_META_INTS + 4], ib = meta[(size_t)b * FR3D_META_INTS + 4];
            const int i1 = min(ia, ib), i2 = max(ia, ib);
            const int x = atomicAdd(&s_nnest, 1);
            nkey[x] = ((tail_u64)(uint32_t)(nt_chain[a] - c_lo) << 48) | ((tail_u64)(uint32_t)(i2 - i1) << 24) | (tail_u64)(uint32_t)i1;
        }
        __syncthreads();
        const int nnest = s_nnest;
        tail_block_sort_keys<TT, tail_u64>(nkey, nnest, tail_smem, ws.tmpk + eb);
        for (int c = c_lo; c < c_hi; ++c) {                                  // identity: every index maps to itself
            const int eo = id.chain_ends_off[c] - ends_lo, len = id.chain_ends_off[c + 1] - id.chain_ends_off[c];
            for (int i = tid; i < len; i += T) ends_base[eo + i] = i;
        }
        __syncthreads();
        // one warp per chain: its candidates are contiguous in the sorted list
        if (nnest > 0) {
            for (int c = warp; c < c_hi - c_lo; c += T / 32) {
                int b0 = 0, b1 = nnest;                              // first candidate with chain >= c
                while (b0 < b1) {
                    const int mid = (b0 + b1) >> 1;
                    if ((int)(nkey[mid] >> 48) < c) b0 = mid + 1; else b1 = mid;
                }
                int32_t *ends = ends_base + (id.chain_ends_off[c_lo + c] - ends_lo);
                for (int x = b0; x < nnest; ++x) {
                    const tail_u64 k = nkey[x];
                    if ((int)(k >> 48) != c) break;
                    const int i1 = (int)(k & 0xFFFFFFull), i2 = i1 + (int)((k >> 24) & 0xFFFFFFull);
                    bool nested = i2 > i1;                           // span 0: never nested (NA:1112-1116)
                    for (int i0 = i1 + 1; i0 < i2 && nested; i0 += 32) {
                        const int i = i0 + lane;
                        bool bad = false;
                        if (i < i2) { const int e = ends[i]; bad = !(e > i1 && e < i2); }
                        if (__any_sync(0xFFFFFFFFu, bad)) nested = false;
                    }
                    if (nested && lane == 0) { ends[i1] = i2; ends[i2] = i1; }
                    __syncwarp();
                }
            }
        }
        __syncthreads();

        TAIL_PHASE(6);
        // ---- crossing numbers (NA:1146-1169) and row counts
        int my_rows = 0;
        for (int k = tid; k < napp; k += T) {
            const int a = aa[k], b = ab[k];
            int crossing = 0;
            const int ca = nt_chain[a];
            if (ca == nt_chain[b]) {
                const int ia = meta[(size_t)a * FR3D_META_INTS + 4], ib = meta[(size_t)b * FR3D_META_INTS + 4];
                const int i1 = min(ia, ib), i2 = max(ia, ib);
                const int32_t *ends = ends_base + (id.chain_ends_off[ca] - ends_lo);
                for (int i = i1 + 1; i < i2; ++i) {
                    const int e = ends[i];
                    crossing += (e < i1 || e > i2);
                }
            }
            across[k] = crossing;
            my_rows += 1 + (int)(tb.labels[alab[k]].flags & 1u);
        }
        TAIL_PHASE(7);
        // ---- label ranks by first appearance, then the rows in append order
        for (int l = tid; l < n_labels; l += T) {
            const uint32_t f = s_lfirst[l];
            int r = 0;
            if (f != TAIL_NONE)
                for (int m = 0; m < n_labels; ++m) r += s_lfirst[m] < f;
            s_lrank[l] = (uint16_t)r;
        }
        __syncthreads();
        // rows ordered by (label rank, append index): the append index is the payload, in the low bits of the key -
        // 32-bit keys (10 + 22 bits) unless the structure has more than four million rows
        const bool keys32 = napp < (1 << 22);
        uint32_t *akey32 = reinterpret_cast<uint32_t *>(akey);
        if (keys32) {
            for (int k = tid; k < napp; k += T) akey32[k] = ((uint32_t)s_lrank[alab[k]] << 22) | (uint32_t)k;
            tail_block_sort_keys<TT, uint32_t>(akey32, napp, tail_smem, ws.tmpk + eb);
        } else {
            for (int k = tid; k < napp; k += T) akey[k] = ((tail_u64)s_lrank[alab[k]] << 32) | (tail_u64)(uint32_t)k;
            tail_block_sort_keys<TT, tail_u64>(akey, napp, tail_smem, ws.tmpk + eb);
        }
#define TAIL_ROW_OF(r) (keys32 ? (int)(akey32[r] & ((1u << 22) - 1u)) : (int)(akey[r] & 0xFFFFFFFFull))

        TAIL_PHASE(8);
        // ---- covalent links (NA:1181-1210): nts by (model symmetry chain, index), successive keys of one string
        tail_u64 *ckey = ws.ckey + lo;
        for (int u = tid; u < n; u += T) {
            const int ix = meta[(size_t)u * FR3D_META_INTS + 4];
            ckey[u] = ((tail_u64)(uint32_t)id.nt_cov[lo + u] << (TAIL_INDEX_BITS + TAIL_MAX_NT_BITS)) |
                      ((tail_u64)(uint32_t)ix << TAIL_MAX_NT_BITS) | (tail_u64)u;
        }
        __syncthreads();
        {       // files usually list the nts of a chain in sequence order: then the keys are sorted already
            int unsorted = 0;
            for (int u = tid; u + 1 < n; u += T) unsorted |= ckey[u] > ckey[u + 1];
            if (__syncthreads_or(unsorted)) tail_block_sort_keys<TT, tail_u64>(ckey, n, tail_smem, ws.tmpc + lo);
        }
#define TAIL_COV_NT(r) ((int)(ckey[r] & ((1ull << TAIL_MAX_NT_BITS) - 1ull)))
        for (int r = tid; r < n; r += T) {                   // rows of sorted position r = members of the next key
            const tail_u64 kr = ckey[r] >> TAIL_MAX_NT_BITS;
            int e = r + 1;
            while (e < n && (ckey[e] >> TAIL_MAX_NT_BITS) == kr) ++e;
            int cnt = 0;
            if (e < n && (ckey[e] >> (TAIL_INDEX_BITS + TAIL_MAX_NT_BITS)) == (kr >> TAIL_INDEX_BITS)) {
                const tail_u64 ke = ckey[e] >> TAIL_MAX_NT_BITS;
                int f = e + 1;
                while (f < n && (ckey[f] >> TAIL_MAX_NT_BITS) == ke) ++f;
                cnt = f - e;
            }
            gcount[r] = cnt;                                 // the per-nt arrays are free again
            gstart[r] = e;
            my_rows += cnt;
        }
        TAIL_PHASE(9);
        int total_rows;
        tail_block_scan<TT>(my_rows, s_w, total_rows);
        __syncthreads();
        if (tid == 0) {
            if (s_err) { atomicOr((tail_u64 *)&tc[2], 2ull); total_rows = 0; }
            const long long b = (long long)atomicAdd((tail_u64 *)&tc[0], (tail_u64)total_rows);
            s_base = b;
            const bool fits = b + total_rows <= (long long)row_cap;
            row_start[s] = fits ? (int32_t)b : -1;
            row_count[s] = s_err ? 0 : total_rows;
            if (!fits) s_err = 1;                            // nothing is written; the host re-runs with the count
        }
        __syncthreads();
        if (s_err) continue;
        fr3d_row_t *out = rows + s_base;
        int running = 0;
        for (int r0 = 0; r0 < napp; r0 += T) {
            const int r = r0 + tid;
            int k = 0, lab = 0, cnt = 0;
            if (r < napp) { k = TAIL_ROW_OF(r); lab = alab[k]; cnt = 1 + (int)(tb.labels[lab].flags & 1u); }
            int total;
            const int ex = tail_block_scan<TT>(cnt, s_w, total);
            if (r < napp) {
                fr3d_row_t row;
                row.label = lab; row.nt1 = aa[k] + base; row.nt2 = ab[k] + base; row.crossing = across[k];
                out[running + ex] = row;
                if (cnt == 2) {                                                          // NA:1171-1177
                    row.label = tb.labels[lab].rev; row.nt1 = ab[k] + base; row.nt2 = aa[k] + base;
                    out[running + ex + 1] = row;
                }
            }
            running += total;
        }
        for (int r0 = 0; r0 < n; r0 += T) {
            const int r = r0 + tid;
            const int cnt = r < n ? gcount[r] : 0;
            int total;
            const int ex = tail_block_scan<TT>(cnt, s_w, total);
            if (cnt) {
                const int u1 = TAIL_COV_NT(r), e = gstart[r];
                const int i1 = meta[(size_t)u1 * FR3D_META_INTS + 4];
                for (int x = 0; x < cnt; ++x) {
                    const int u2 = TAIL_COV_NT(e + x);
                    fr3d_row_t row;
                    row.label = FR3D_LABEL_COVALENT | (meta[(size_t)u2 * FR3D_META_INTS + 4] - i1);
                    row.nt1 = u1 + base; row.nt2 = u2 + base; row.crossing = -1;
                    out[running + ex + x] = row;
                }
            }
            running += total;
        }
        TAIL_PHASE(10);
    }
#undef TAIL_HIT_OF
#undef TAIL_ROW_OF
#undef TAIL_COV_NT
}

}  // namespace fr3d
