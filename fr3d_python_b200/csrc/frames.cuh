// K1: per-nucleotide base frames + hydrogen placement, one thread per nt.
//
// Replaces Component.calculate_rotation_matrix (fr3d/data/components.py:273-349)
// -> besttransformation (fr3d/geometry/superpositions.py:10-89) and the
// arithmetic of infer_NA_hydrogens (components.py:452-462).
//
// Work per nt is tiny (<= 11 atoms): one thread per nt keeps the 4x4 Jacobi in registers.  The 32 records of a
// warp are one contiguous 12 KB block of base_xyz, so the warp first copies it into shared memory with coalesced
// cp.async (a thread reading "its" record from global memory makes every load touch 32 cache lines) and each
// thread then works in its own row (odd stride: no bank conflicts); the slots a thread may have written
// (hydrogens, ideal-position copies) are copied back the same way.  The standard base coordinates are read
// with a per-lane parent index, which would serialise constant-cache loads: they sit in shared memory too.
#pragma once

#include "classify_pm.cuh"
#include "rotfit.cuh"
#include "tables.cuh"

namespace fr3d {

constexpr int FRAMES_WARPS = 3;
constexpr int FRAMES_STRIDE = FR3D_BASE_SLOTS * 3 + 1;     // 49 doubles per row

// the fit of one nt; xyz = its 16 x 3 slot block (a shared-memory row), std_xyz = the standard base coordinates.
// Returns true if a slot below 8 was written (ideal-position copies of a modified residue).
// mask_in: the observed-atom mask the fit starts from - the record's own base_mask, or (atoms read from a compact
// plane) the uploaded mask plane, which is never written: the fit is then a pure function of the upload and can be
// repeated on records a first pass has already completed.
__device__ __forceinline__ bool frames_fit_one(const fr3d_nts_t &nts, int32_t *__restrict__ status, int place_h, int n,
                                               double *xyz, const double (*std_xyz)[FR3D_BASE_SLOTS][3],
                                               uint32_t mask_in) {
    bool wrote_low = false;
    const int parent = nts.meta[n * FR3D_META_INTS + 0];
    const int flags = nts.meta[n * FR3D_META_INTS + 1];
    uint32_t mask = mask_in & 0xFFFFu;
    // residues that already received their ideal-position copies hold averaged coordinates: fitting
    // them again would not reproduce the frame, so a second pass leaves them alone
    if ((flags & FR3D_FLAG_DUP) && (mask_in & FR3D_MASK_HAS_FRAME)) {
        if (status) status[n] = 0;
        return false;
    }
    // bits 16-19: parent + 1, bits 20-23: flags (read by the classify kernel instead of meta)
    const uint32_t tag = ((uint32_t)(parent + 1) & 15u) << 16 | ((uint32_t)flags & 15u) << 20;
    if (parent < 0 || parent >= FR3D_N_PARENTS) {
        nts.base_mask[n] = mask;
        if (status) status[n] = -1;
        return false;
    }
    const int nh = c_tab.n_heavy[parent];
    const int ns = c_tab.n_slots[parent];

    // means over the observed heavy atoms (superpositions.py:40-41)
    double mr[3] = {0, 0, 0}, ms[3] = {0, 0, 0};
    int cnt = 0;
    for (int k = 0; k < nh; ++k) {
        if (mask >> k & 1u) {
            mr[0] += xyz[k * 3 + 0]; mr[1] += xyz[k * 3 + 1]; mr[2] += xyz[k * 3 + 2];
            ms[0] += std_xyz[parent][k][0]; ms[1] += std_xyz[parent][k][1];
            ms[2] += std_xyz[parent][k][2];
            ++cnt;
        }
    }
    if (cnt < 3) {
        // fewer than three base atoms: LAPACK's null-space choice is not reproducible
        // (SURVEY §7 hard part 5); the nt is left without a frame, like a failed fit
        // (components.py:324-337), and drops out of the cubes (NA:390,426).
        nts.base_mask[n] = mask | tag;
        if (status) status[n] = -2;
        return false;
    }
    const double inv = 1.0 / (double)cnt;
    mr[0] *= inv; mr[1] *= inv; mr[2] *= inv;
    ms[0] *= inv; ms[1] *= inv; ms[2] *= inv;

    // M[i][j] = sum_k dev1_k[i] * dev2_k[j]  (= A^T of superpositions.py:53)
    double M[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    double ssq = 0.0;                                      // sum |a|^2 + sum |b|^2: twice the eigenvalue bound
    for (int k = 0; k < nh; ++k) {
        if (mask >> k & 1u) {
            const double a0 = xyz[k * 3 + 0] - mr[0], a1 = xyz[k * 3 + 1] - mr[1], a2 = xyz[k * 3 + 2] - mr[2];
            const double b0 = std_xyz[parent][k][0] - ms[0], b1 = std_xyz[parent][k][1] - ms[1],
                         b2 = std_xyz[parent][k][2] - ms[2];
            M[0][0] += a0 * b0; M[0][1] += a0 * b1; M[0][2] += a0 * b2;
            M[1][0] += a1 * b0; M[1][1] += a1 * b1; M[1][2] += a1 * b2;
            M[2][0] += a2 * b0; M[2][1] += a2 * b1; M[2][2] += a2 * b2;
            ssq += (a0 * a0 + a1 * a1 + a2 * a2) + (b0 * b0 + b1 * b1 + b2 * b2);
        }
    }
    double Q[3][3];
#ifdef FR3D_FRAMES_JACOBI
    horn_rotation(M, Q);
#else
    // largest eigenvector of Horn's matrix from its characteristic quartic + a pivoted adjugate column (rotfit.cuh):
    // the cyclic Jacobi sweeps were 55 % of this kernel's instructions; both agree with numpy's SVD route to 1e-14
    horn_rotation_pivot(M, 0.5 * ssq, Q);
#endif
    // dev1 . U ~ dev2 with row vectors  <=>  U = Q^T
    double U[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) U[r][c] = Q[c][r];

    double ctr[3] = {mr[0], mr[1], mr[2]};
    if (flags & FR3D_FLAG_MODIFIED) {   // components.py:348-349
#pragma unroll
        for (int r = 0; r < 3; ++r) ctr[r] = mr[r] - (U[r][0] * ms[0] + U[r][1] * ms[1] + U[r][2] * ms[2]);
    }
    double *co = nts.center + (size_t)n * 3;
    co[0] = ctr[0]; co[1] = ctr[1]; co[2] = ctr[2];
    double *ro = nts.rot + (size_t)n * 9;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) ro[r * 3 + c] = U[r][c];

    if (flags & FR3D_FLAG_DUP) {
        // components.py:505-513: an ideal-position copy of every mapped base atom (heavy atoms included) is
        // appended; AtomProxy then returns the mean of the observed atom and its copy (base.py:150-166)
        const uint32_t dup = (uint32_t)nts.meta[n * FR3D_META_INTS + 7] & 0xFFFFu;
        uint32_t twin = 0;
        for (int k = 0; k < ns; ++k) {
            if (dup >> k & 1u) {
                const double h0 = std_xyz[parent][k][0], h1 = std_xyz[parent][k][1],
                             h2 = std_xyz[parent][k][2];
                const double i0 = ctr[0] + (U[0][0] * h0 + U[0][1] * h1 + U[0][2] * h2);
                const double i1 = ctr[1] + (U[1][0] * h0 + U[1][1] * h1 + U[1][2] * h2);
                const double i2 = ctr[2] + (U[2][0] * h0 + U[2][1] * h1 + U[2][2] * h2);
                if (k < 8) wrote_low = true;
                if (mask >> k & 1u) {
                    xyz[k * 3 + 0] = (xyz[k * 3 + 0] + i0) / 2.0;
                    xyz[k * 3 + 1] = (xyz[k * 3 + 1] + i1) / 2.0;
                    xyz[k * 3 + 2] = (xyz[k * 3 + 2] + i2) / 2.0;
                    twin |= 1u << k;
                } else {
                    xyz[k * 3 + 0] = i0; xyz[k * 3 + 1] = i1; xyz[k * 3 + 2] = i2;
                    mask |= 1u << k;
                }
            }
        }
        nts.meta[n * FR3D_META_INTS + 7] = (int32_t)(dup | (twin << 16));
    } else if (place_h && !(flags & FR3D_FLAG_MODIFIED)) {
        // components.py:452-462: centre + U . h_std for every missing base hydrogen of a standard residue
        for (int k = nh; k < ns; ++k) {
            if (!(mask >> k & 1u)) {
                const double h0 = std_xyz[parent][k][0], h1 = std_xyz[parent][k][1],
                             h2 = std_xyz[parent][k][2];
                xyz[k * 3 + 0] = ctr[0] + (U[0][0] * h0 + U[0][1] * h1 + U[0][2] * h2);
                xyz[k * 3 + 1] = ctr[1] + (U[1][0] * h0 + U[1][1] * h1 + U[1][2] * h2);
                xyz[k * 3 + 2] = ctr[2] + (U[2][0] * h0 + U[2][1] * h1 + U[2][2] * h2);
                mask |= 1u << k;
            }
        }
    }
    nts.base_mask[n] = mask | tag | FR3D_MASK_HAS_FRAME;
    if (status) status[n] = 0;
    return wrote_low;
}

// src_plane / src_cols / src_mask: where the observed atoms are read from - nts.base_xyz and nts.base_mask themselves
// (nullptr), or a compact plane [n][src_cols / 3][3] holding only the first slots of every record (the upload keeps
// slots 0..10 when no hydrogen sits beyond them: they are placed here anyway) with its own mask plane; the full
// 16-slot record and the completed mask are then written to nts.base_xyz / nts.base_mask.
//
// group_off != nullptr: src_plane is RAGGED - only the observed atoms of slots 0..10, nt by nt in slot order, and
// group_off[g] is the first atom of the 32 nts of warp g (the upload then carries ~9.3 instead of 11 slots per nt).  The
// warp copies its contiguous piece into the far end of its row buffer, every lane takes its own atoms (offset = a warp scan
// of the mask bit counts) into registers, and after a __syncwarp writes them to their slots of its row.
#ifndef FR3D_FRAMES_MINB
#define FR3D_FRAMES_MINB 5      // 128 registers, five blocks per SM (the shared-memory limit): 87.7 -> 84.4 us per 316 k nts
#endif
__global__ void __launch_bounds__(FRAMES_WARPS * 32, FR3D_FRAMES_MINB) frames_fit_kernel(fr3d_nts_t nts, int32_t *__restrict__ status,
                                                                       int place_h, const double *__restrict__ src_plane,
                                                                       int src_cols, const uint32_t *__restrict__ src_mask,
                                                                       const int32_t *__restrict__ group_off) {
    __shared__ double s_std[FR3D_N_PARENTS][FR3D_BASE_SLOTS][3];
    __shared__ double s_rows[FRAMES_WARPS][32 * FRAMES_STRIDE];
    for (int item = threadIdx.x; item < FR3D_N_PARENTS * FR3D_BASE_SLOTS * 3; item += blockDim.x)
        (&s_std[0][0][0])[item] = (&c_tab.std_xyz[0][0][0])[item];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n0 = (blockIdx.x * FRAMES_WARPS + warp) * 32;              // first nt of this warp
    const int n = n0 + lane;
    double *rows = s_rows[warp];
    const int rows_here = min(32, max(0, nts.n_nt - n0));
    const bool compact = src_plane != nullptr;
    const int cols = compact ? src_cols : FR3D_BASE_SLOTS * 3;
    if (group_off) {
        constexpr int HEAVY = 11;                                         // every heavy base atom lies in slots 0..10
        const int g = blockIdx.x * FRAMES_WARPS + warp;
        double *flat = rows + 32 * FRAMES_STRIDE - 32 * HEAVY * 3;        // the last 1056 doubles of the warp's buffer
        uint32_t m = 0;
        int excl = 0;
        if (rows_here > 0) {
            const int a0 = group_off[g], a1 = group_off[g + 1];
            for (int t = lane; t < (a1 - a0) * 3; t += 32) cp_async8(flat + t, src_plane + (size_t)a0 * 3 + t);
            m = n < nts.n_nt ? (src_mask[n] & ((1u << HEAVY) - 1u)) : 0u;
            int incl = __popc(m);
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int v = __shfl_up_sync(0xFFFFFFFFu, incl, off);
                if (lane >= off) incl += v;
            }
            excl = incl - __popc(m);
            cp_async_wait_all();
        }
        __syncwarp();
        double v[HEAVY][3];
        const int cnt = __popc(m);
#pragma unroll
        for (int k = 0; k < HEAVY; ++k) {
            if (k < cnt) {
                v[k][0] = flat[(excl + k) * 3]; v[k][1] = flat[(excl + k) * 3 + 1]; v[k][2] = flat[(excl + k) * 3 + 2];
            }
        }
        __syncwarp();                                                     // everybody has its atoms: the rows may be written
        double *row = rows + lane * FRAMES_STRIDE;
        for (int c = 0; c < FR3D_BASE_SLOTS * 3; ++c) row[c] = 0.0;
#pragma unroll
        for (int k = 0; k < HEAVY; ++k) {
            if (k < cnt) {
                const int slot = __fns(m, 0, k + 1);                      // the (k + 1)-th observed slot
                row[slot * 3] = v[k][0]; row[slot * 3 + 1] = v[k][1]; row[slot * 3 + 2] = v[k][2];
            }
        }
    } else
    {   // stage the warp's records row by row (cols doubles each: lanes 0..31, then the first lanes for the rest)
        const double *src = (compact ? src_plane : nts.base_xyz) + (size_t)n0 * cols + lane;
        double *dst = rows + lane;
#pragma unroll 4
        for (int r = 0; r < rows_here; ++r) {
            if (lane < cols) cp_async8(dst + r * FRAMES_STRIDE, src + (size_t)r * cols);
            if (lane < cols - 32) cp_async8(dst + r * FRAMES_STRIDE + 32, src + (size_t)r * cols + 32);
            if (compact && lane >= cols - 32 && lane < FR3D_BASE_SLOTS * 3 - 32) dst[r * FRAMES_STRIDE + 32] = 0.0;
        }
        cp_async_wait_all();
    }
    __syncthreads();
    bool wrote_low = false;                                               // wrote a slot below 8 (ideal-position copies)
    if (n < nts.n_nt)
        wrote_low = frames_fit_one(nts, status, place_h, n, rows + lane * FRAMES_STRIDE, s_std,
                                   src_mask ? src_mask[n] : nts.base_mask[n]);
    __syncwarp();
    // copy back slots 8..15 (every hydrogen slot lies there: the parents have 8 to 11 heavy atoms), or whole rows
    // if some lane inserted ideal-position copies or the records came from a compact plane
    double *dstg = nts.base_xyz + (size_t)n0 * FR3D_BASE_SLOTS * 3;
    if (compact || __any_sync(0xFFFFFFFFu, wrote_low)) {
        for (int r = 0; r < rows_here; ++r) {
            dstg[r * (FR3D_BASE_SLOTS * 3) + lane] = rows[r * FRAMES_STRIDE + lane];
            if (lane < FR3D_BASE_SLOTS * 3 - 32) dstg[r * (FR3D_BASE_SLOTS * 3) + 32 + lane] = rows[r * FRAMES_STRIDE + 32 + lane];
        }
    } else {        // columns 24..47: 24 doubles per row, lanes 0..23
#pragma unroll 4
        for (int r = 0; r < rows_here; ++r)
            if (lane < 24) dstg[r * (FR3D_BASE_SLOTS * 3) + 24 + lane] = rows[r * FRAMES_STRIDE + 24 + lane];
    }
}

// The upload's backbone plane carries slots 0..8; slot 9 (the O3' of the previous residue, NA:480-525) is filled here
// from the previous ROW where prev_rel says so (the common case) and from a short explicit list otherwise.
__global__ void backbone_expand_kernel(const double *__restrict__ bb9, const int8_t *__restrict__ prev_rel, int n_nt,
                                       const int32_t *__restrict__ extra_nt, const double *__restrict__ extra_xyz, int n_extra,
                                       double *__restrict__ bb_xyz) {
    const long long total = (long long)n_nt * FR3D_BB_SLOTS * 3;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(t / (FR3D_BB_SLOTS * 3)), c = (int)(t % (FR3D_BB_SLOTS * 3));
        double v = 0.0;
        if (c < 27) v = bb9[(size_t)n * 27 + c];
        else if (prev_rel[n] == 1 && n > 0) v = bb9[(size_t)(n - 1) * 27 + 5 * 3 + (c - 27)];
        bb_xyz[t] = v;
    }
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_extra * 3; e += gridDim.x * blockDim.x)
        bb_xyz[(size_t)extra_nt[e / 3] * FR3D_BB_SLOTS * 3 + 27 + e % 3] = extra_xyz[e];
}

}  // namespace fr3d
