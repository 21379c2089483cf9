// libfr3d_b200.so — C ABI over the sm_100a kernels (see include/fr3d_b200.h).
// Single translation unit: the kernels share one __constant__ table block.

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/fr3d_b200.h"
#include "tables.cuh"
#include "rotfit.cuh"
#include "frames.cuh"
#include "search.cuh"
#include "classify.cuh"
#include "classify_pm.cuh"
#include "classify_tpp.cuh"
#include "classify_screen.cuh"
#include "classify_screen32.cuh"
#include "classify_list2.cuh"
#include "classify_stack32.cuh"
#include "geometry.cuh"
#include "unit_annotation.cuh"
#include "radius_search.cuh"
#include "tail.cuh"
#include "ordering.cuh"
#include "possibilities.cuh"

#ifndef FR3D_BPH_MINB
#define FR3D_BPH_MINB 6      // base-phosphate list kernel, min blocks per SM 3 / 4 / 6 / 8: classify 0.840 / 0.835 / 0.807 / 0.815 ms
#endif
#ifndef FR3D_LIGHT_MINB
#define FR3D_LIGHT_MINB 4
#endif

namespace {

thread_local char g_err[512] = "";

// Per-device state: the __constant__ tables and the shared-memory opt-ins live per device, so "initialised" and
// "tables uploaded" are tracked per device too (one thread per GPU in one process is a supported way to drive this
// library).  Written under g_mu; read without it after the owning thread's fr3d_init returned.
constexpr int MAX_DEVICES = 64;
struct DeviceState {
    bool initialised = false;
    bool tables_loaded = false;
    int sm_count = 148;
};
DeviceState g_dev[MAX_DEVICES];
std::mutex g_mu;

DeviceState *current_device_state() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= MAX_DEVICES) return nullptr;
    return &g_dev[d];
}

// Tuning knobs (environment), read once per process:
// FR3D_CLASSIFY_MODE 2 = screen + work-list kernels (default when a workspace is given), 1 = thread-per-pair kernel
// over all pairs, 0 = warp-per-pair kernel over all pairs; FR3D_SCREEN=64 / FR3D_STACK=64: the fp64 screen / stacking
// kernels (A/B runs); FR3D_CLASSIFY_GRID_MULT: blocks per SM.
struct Knobs {
    int mult = 0, mode = 2, screen64 = 0, stack64 = 0;
};
Knobs g_knobs;
std::once_flag g_knobs_once;
const Knobs &knobs() {
    std::call_once(g_knobs_once, [] {
        const char *s64 = getenv("FR3D_SCREEN");
        g_knobs.screen64 = s64 && atoi(s64) == 64;
        const char *t64 = getenv("FR3D_STACK");
        g_knobs.stack64 = g_knobs.screen64 || (t64 && atoi(t64) == 64);   // the fp32 stacking kernel reads the screen records
        const char *e = getenv("FR3D_CLASSIFY_GRID_MULT");
        g_knobs.mult = (e && atoi(e) > 0) ? atoi(e) : 0;
        const char *m = getenv("FR3D_CLASSIFY_MODE");
        g_knobs.mode = m ? atoi(m) : 2;
    });
    return g_knobs;
}

int fail(int code, const char *what, cudaError_t e = cudaSuccess) {
    if (e != cudaSuccess) snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    else snprintf(g_err, sizeof(g_err), "%s", what);
    return code;
}

#define CUDA_TRY(expr)                                                 \
    do {                                                               \
        cudaError_t _e = (expr);                                       \
        if (_e != cudaSuccess) return fail(FR3D_E_CUDA, #expr, _e);    \
    } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

uint32_t table_size_for(int32_t n_nt) {
    uint32_t t = 64;
    while (t < (uint32_t)n_nt * 2u) t <<= 1;
    return t;
}

// carve the caller's workspace into the search arrays
fr3d::SearchWs carve(void *workspace, int32_t n_nt) {
    fr3d::SearchWs ws;
    const uint32_t T = table_size_for(n_nt);
    char *p = static_cast<char *>(workspace);
    ws.keys = reinterpret_cast<unsigned long long *>(p); p += align_up(sizeof(unsigned long long) * T, 256);
    ws.count = reinterpret_cast<int32_t *>(p); p += align_up(sizeof(int32_t) * T, 256);
    ws.first = reinterpret_cast<int32_t *>(p); p += align_up(sizeof(int32_t) * T, 256);
    ws.start = reinterpret_cast<int32_t *>(p); p += align_up(sizeof(int32_t) * T, 256);
    ws.cursor = reinterpret_cast<int32_t *>(p); p += align_up(sizeof(int32_t) * T, 256);
    ws.members = reinterpret_cast<int32_t *>(p); p += align_up(sizeof(int32_t) * (size_t)n_nt, 256);
    ws.cell_slot = reinterpret_cast<int32_t *>(p); p += align_up(sizeof(int32_t) * (size_t)n_nt, 256);
    ws.cell_xyz = reinterpret_cast<int32_t *>(p); p += align_up(sizeof(int32_t) * 3 * (size_t)n_nt, 256);
    ws.cell_list = reinterpret_cast<int32_t *>(p); p += align_up(sizeof(int32_t) * (size_t)n_nt, 256);
    ws.n_cells = reinterpret_cast<int32_t *>(p); p += 256;
    ws.cell_info = reinterpret_cast<int4 *>(p); p += align_up(sizeof(int4) * T, 256);
    ws.sorted = reinterpret_cast<double2 *>(p); p += align_up(sizeof(double2) * 3 * (size_t)n_nt, 256);
    ws.table_mask = T - 1;
    return ws;
}

// rank = first nt * 16 + stencil (+ 16 * index_offset) is int32: nt indices must stay below 2^27
constexpr long long MAX_NT_INDEX = 1ll << 27;

}  // namespace

extern "C" {

int fr3d_version(void) { return FR3D_ABI_VERSION; }

const char *fr3d_last_error(void) { return g_err; }

static int init_current_device(int device) {
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(FR3D_E_CUDA, "libfr3d_b200 is built for sm_100a only");
    g_dev[device].sm_count = prop.multiProcessorCount;
    // opt in to the large dynamic shared-memory carve-outs once per device (not on every call)
    CUDA_TRY(cudaFuncSetAttribute(fr3d::classify_screen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)fr3d::SCR_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(fr3d::classify_screen32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)fr3d::S32_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(fr3d::classify_stack32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)fr3d::ST32_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(fr3d::classify_list2_kernel<false, fr3d::L2_WARPS_STACK>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)fr3d::l2_smem_bytes<false, fr3d::L2_WARPS_STACK>()));
    CUDA_TRY(cudaFuncSetAttribute(fr3d::classify_list2_kernel<true, fr3d::L2_WARPS_PAIR>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)fr3d::l2_smem_bytes<true, fr3d::L2_WARPS_PAIR>()));
    CUDA_TRY(cudaFuncSetAttribute(fr3d::kabsch_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)fr3d::KABSCH_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(fr3d::discrepancy_all_vs_all_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)fr3d::disc_smem_bytes(fr3d::DISC_MAXN)));
    CUDA_TRY(cudaFuncSetAttribute(fr3d::tail_structure_kernel<fr3d::TAIL_THREADS, FR3D_TAIL_MINB, false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fr3d::TAIL_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(fr3d::tail_structure_kernel<fr3d::TAIL_THREADS_MANY, 8, false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)fr3d::tail_smem_bytes<fr3d::TAIL_THREADS_MANY>()));
    CUDA_TRY(cudaFuncSetAttribute(fr3d::tail_structure_kernel<fr3d::TAIL_THREADS_BIG, 1, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)fr3d::tail_smem_bytes<fr3d::TAIL_THREADS_BIG>()));
    g_dev[device].initialised = true;
    return FR3D_OK;
}

int fr3d_init(int device) {
    if (device < 0 || device >= MAX_DEVICES) return fail(FR3D_E_ARG, "device out of range");
    std::lock_guard<std::mutex> lock(g_mu);
    int previous = -1;
    CUDA_TRY(cudaGetDevice(&previous));          // the caller's current device is left as it was
    CUDA_TRY(cudaSetDevice(device));
    const int rc = init_current_device(device);
    if (previous >= 0 && previous != device) cudaSetDevice(previous);
    knobs();
    return rc;
}

int fr3d_upload_tables(const fr3d_static_tables_t *host_tables) {
    if (!host_tables) return fail(FR3D_E_ARG, "host_tables is NULL");
    DeviceState *ds = current_device_state();
    if (!ds || !ds->initialised) return fail(FR3D_E_ARG, "fr3d_init has not been called for the current device");
    std::lock_guard<std::mutex> lock(g_mu);
    CUDA_TRY(cudaMemcpyToSymbol(fr3d::c_tab, host_tables, sizeof(fr3d_static_tables_t)));
    ds->tables_loaded = true;
    return FR3D_OK;
}

// every compute entry point: the current device must have been initialised (and, where rule tables are read, loaded)
static int require_device(bool tables, int *sm_count) {
    DeviceState *ds = current_device_state();
    if (!ds || !ds->initialised) return fail(FR3D_E_ARG, "fr3d_init has not been called for the current device");
    if (tables && !ds->tables_loaded) return fail(FR3D_E_NOTABLES, "fr3d_upload_tables not called on the current device");
    if (sm_count) *sm_count = ds->sm_count;
    return FR3D_OK;
}

int fr3d_frames_fit_from(const fr3d_nts_t *nts, const double *compact_base_xyz, int32_t compact_slots,
                         const uint32_t *observed_mask, int32_t *status, int place_h, void *stream) {
    if (!nts) return fail(FR3D_E_ARG, "nts is NULL");
    if (compact_base_xyz && (compact_slots < 11 || compact_slots > FR3D_BASE_SLOTS))
        return fail(FR3D_E_ARG, "compact_slots must be in [11, 16] (every heavy base atom lies in slots 0..10)");
    if (compact_base_xyz && (!observed_mask || observed_mask == nts->base_mask))
        return fail(FR3D_E_ARG, "a compact plane needs its own observed_mask plane (not nts->base_mask, which is written)");
    if (!compact_base_xyz) observed_mask = nullptr;
    if (int rc = require_device(true, nullptr)) return rc;
    if (nts->n_nt <= 0) return FR3D_OK;
    const int block = fr3d::FRAMES_WARPS * 32;
    const int grid = (nts->n_nt + block - 1) / block;
    fr3d::frames_fit_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(*nts, status, place_h, compact_base_xyz,
                                                                                   compact_slots * 3, observed_mask, nullptr);
    CUDA_TRY(cudaGetLastError());
    return FR3D_OK;
}

int fr3d_frames_fit_ragged(const fr3d_nts_t *nts, const double *ragged_xyz, const int32_t *group_off,
                           const uint32_t *observed_mask, int32_t *status, int place_h, void *stream) {
    if (!nts || !ragged_xyz || !group_off || !observed_mask) return fail(FR3D_E_ARG, "NULL argument");
    if (observed_mask == nts->base_mask) return fail(FR3D_E_ARG, "the observed_mask plane must not be nts->base_mask (which is written)");
    if (int rc = require_device(true, nullptr)) return rc;
    if (nts->n_nt <= 0) return FR3D_OK;
    const int block = fr3d::FRAMES_WARPS * 32;
    const int grid = (nts->n_nt + block - 1) / block;
    fr3d::frames_fit_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(*nts, status, place_h, ragged_xyz, 33,
                                                                                   observed_mask, group_off);
    CUDA_TRY(cudaGetLastError());
    return FR3D_OK;
}

int fr3d_backbone_expand(const double *bb9, const int8_t *prev_rel, int32_t n_nt, const int32_t *extra_nt,
                         const double *extra_xyz, int32_t n_extra, double *bb_xyz, void *stream) {
    if (!bb9 || !prev_rel || !bb_xyz || (n_extra > 0 && (!extra_nt || !extra_xyz))) return fail(FR3D_E_ARG, "NULL argument");
    if (n_nt < 0 || n_extra < 0) return fail(FR3D_E_ARG, "negative size");
    int sm_count = 148;
    if (int rc = require_device(false, &sm_count)) return rc;
    if (n_nt == 0) return FR3D_OK;
    const long long want = ((long long)n_nt * FR3D_BB_SLOTS * 3 + 255) / 256;
    fr3d::backbone_expand_kernel<<<(int)(want < sm_count * 8 ? want : sm_count * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        bb9, prev_rel, n_nt, extra_nt, extra_xyz, n_extra, bb_xyz);
    CUDA_TRY(cudaGetLastError());
    return FR3D_OK;
}

int fr3d_frames_fit(const fr3d_nts_t *nts, int32_t *status, int place_h, void *stream) {
    return fr3d_frames_fit_from(nts, nullptr, FR3D_BASE_SLOTS, nullptr, status, place_h, stream);
}

size_t fr3d_pair_search_workspace(int32_t n_nt) {
    if (n_nt < 0) n_nt = 0;
    const uint32_t T = table_size_for(n_nt);
    size_t b = 0;
    b += align_up(sizeof(unsigned long long) * T, 256);
    b += 4 * align_up(sizeof(int32_t) * T, 256);
    b += 3 * align_up(sizeof(int32_t) * (size_t)n_nt, 256);
    b += align_up(sizeof(int32_t) * 3 * (size_t)n_nt, 256);
    b += align_up(sizeof(int4) * T, 256);                   // cell_info
    b += align_up(sizeof(double2) * 3 * (size_t)n_nt, 256); // sorted
    return b + 512;
}

int fr3d_pair_search(const fr3d_nts_t *nts, double cell_size, void *workspace, size_t workspace_bytes,
                     int32_t *pair_i, int32_t *pair_j, int32_t *pair_rank, int64_t pair_cap, int64_t *counters,
                     void *stream) {
    if (!nts || !workspace || !pair_i || !pair_j || !pair_rank || !counters) return fail(FR3D_E_ARG, "NULL argument");
    if (!(cell_size > 0)) return fail(FR3D_E_ARG, "cell_size must be positive");
    if (workspace_bytes < fr3d_pair_search_workspace(nts->n_nt)) return fail(FR3D_E_ARG, "workspace too small");
    if ((long long)nts->n_nt >= MAX_NT_INDEX) return fail(FR3D_E_ARG, "more than 2^27 nts in one call (the visit-order rank is int32)");
    int sm_count = 148;
    if (int rc = require_device(false, &sm_count)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaMemsetAsync(counters, 0, 8 * sizeof(int64_t), st));
    if (nts->n_nt <= 0) return FR3D_OK;
    fr3d::SearchWs ws = carve(workspace, nts->n_nt);
    const uint32_t T = ws.table_mask + 1;
    const int block = 256;
    fr3d::search_reset_kernel<<<(T + block - 1) / block, block, 0, st>>>(ws);
    fr3d::cell_assign_kernel<<<(nts->n_nt + block - 1) / block, block, 0, st>>>(*nts, ws, cell_size, counters);
    fr3d::cell_offsets_kernel<<<(T + fr3d::CELL_OFFSETS_THREADS - 1) / fr3d::CELL_OFFSETS_THREADS, fr3d::CELL_OFFSETS_THREADS, 0, st>>>(ws, counters);
    fr3d::cell_fill_kernel<<<(nts->n_nt + block - 1) / block, block, 0, st>>>(*nts, ws);
    // one warp per occupied cube; the cube count lives on the device, so size the grid for the worst case
    // (a cube per nt) capped at a few resident waves - the kernel strides over the cube list
    const long long want = ((long long)nts->n_nt + fr3d::SEARCH_WARPS - 1) / fr3d::SEARCH_WARPS;
    const long long cap_blocks = (long long)sm_count * 32;
    fr3d::pair_search_kernel<<<(int)(want < cap_blocks ? want : cap_blocks), fr3d::SEARCH_WARPS * 32, 0, st>>>(
        *nts, ws, cell_size, pair_i, pair_j, pair_rank, pair_cap, counters);
    CUDA_TRY(cudaGetLastError());
    return FR3D_OK;
}

size_t fr3d_classify_workspace(int64_t pair_cap, int32_t n_nt) {
    if (pair_cap < 0) pair_cap = 0;
    if (n_nt < 0) n_nt = 0;
    return 6 * align_up(sizeof(int32_t) * (size_t)pair_cap, 256) +
           align_up(sizeof(float) * fr3d::S32_REC * (size_t)n_nt, 256) + 256;
}

int fr3d_classify_pairs(const fr3d_nts_t *nts, const fr3d_box_t *boxes, const int32_t *box_index,
                        const int32_t *pair_i, const int32_t *pair_j, const int32_t *pair_rank, int64_t pair_cap,
                        uint32_t category_mask, int32_t index_offset, void *workspace, size_t workspace_bytes,
                        fr3d_hit_t *hits, int64_t hit_cap, int64_t *counters, void *stream) {
    if (!nts || !boxes || !box_index || !pair_i || !pair_j || !pair_rank || !hits || !counters)
        return fail(FR3D_E_ARG, "NULL argument");
    int sm_count = 148;
    if (int rc = require_device(true, &sm_count)) return rc;
    if (pair_cap <= 0) return FR3D_OK;
    if (workspace && (reinterpret_cast<size_t>(workspace) & 15)) return fail(FR3D_E_ARG, "workspace must be 16-byte aligned");
    if (index_offset < 0 || (long long)nts->n_nt + index_offset >= MAX_NT_INDEX)
        return fail(FR3D_E_ARG, "nt index (with index_offset) reaches 2^27 (the visit-order rank is int32)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Knobs &kn = knobs();
    const int mult = kn.mult, mode = kn.mode, screen64 = kn.screen64, stack64 = kn.stack64;
    int use_mode = mode;
    if (use_mode == 2 && (!workspace || workspace_bytes < fr3d_classify_workspace(pair_cap, nts->n_nt))) use_mode = 0;
    const long long tpp_grid_max = (long long)sm_count * (mult ? mult : 4 * FR3D_TPP_MIN_BLOCKS);
    const long long pm_grid_max = (long long)sm_count * (mult ? mult : FR3D_CLASSIFY_MIN_BLOCKS);
    const long long per_block = (long long)fr3d::WARPS_PER_BLOCK * fr3d::PM_K;
    auto clampg = [](long long want, long long mx) { long long g = want < mx ? want : mx; return (int)(g < 1 ? 1 : g); };
    if (use_mode == 2) {
        // work lists: counters[4] light detail (sO, sugar-ribose), [5] stacking, [6] pairing, [7] base-phosphate
        const size_t list_bytes = align_up(sizeof(int32_t) * (size_t)pair_cap, 256);
        char *wsp = static_cast<char *>(workspace);
        int32_t *wl_light = reinterpret_cast<int32_t *>(wsp);
        int32_t *wl_stack = reinterpret_cast<int32_t *>(wsp + list_bytes);
        int32_t *wl_pair = reinterpret_cast<int32_t *>(wsp + 2 * list_bytes);
        int32_t *wl_bph = reinterpret_cast<int32_t *>(wsp + 3 * list_bytes);
        // [4] undecided stacking pairs, [5] candidate (site, oxygen) combinations of the base-phosphate list
        uint32_t *wl_bph_combos = reinterpret_cast<uint32_t *>(wsp + 5 * list_bytes);
        CUDA_TRY(cudaMemsetAsync(counters + 4, 0, 4 * sizeof(int64_t), st));
        if (screen64) {          // the double-precision screen, kept for A/B runs
            const int scr_threads = fr3d::SCR_WARPS * 32;
            fr3d::classify_screen_kernel<<<clampg((pair_cap + scr_threads - 1) / scr_threads,
                                                  (long long)sm_count * (mult ? mult : 1)),
                                           scr_threads, fr3d::SCR_SMEM, st>>>(
                *nts, boxes, box_index, pair_i, pair_j, pair_cap, category_mask, wl_light, wl_stack, wl_pair, wl_bph, counters);
        } else {
            // single-precision screen records (classify_screen32.cuh), then the screen itself on them
            float *srec = reinterpret_cast<float *>(wsp + 6 * list_bytes);
            if (nts->n_nt > 0)
                fr3d::screen_records_kernel<<<(nts->n_nt + fr3d::SREC_NTS - 1) / fr3d::SREC_NTS, fr3d::SREC_THREADS, 0, st>>>(
                *nts, srec, counters);
            const int scr_threads = fr3d::S32_WARPS * 32;
            fr3d::classify_screen32_kernel<<<clampg((pair_cap + scr_threads - 1) / scr_threads,
                                                    (long long)sm_count * (mult ? mult : 1)),
                                             scr_threads, fr3d::S32_SMEM, st>>>(
                srec, boxes, box_index, pair_i, pair_j, pair_cap, category_mask, wl_light, wl_stack, wl_pair, wl_bph,
                wl_bph_combos, counters);
        }
        // one thread-per-pair launch per list, each instantiation compiled with its own rule families only
        const int tpp_grid = clampg((pair_cap + 127) / 128, tpp_grid_max);
        constexpr uint32_t LIGHT_CATS = FR3D_CAT_SO | FR3D_CAT_SUGAR_RIBOSE;
        if (category_mask & LIGHT_CATS)
            fr3d::classify_pairs_tpp_kernel<FR3D_LIGHT_MINB, LIGHT_CATS, false><<<tpp_grid, 128, 0, st>>>(
                *nts, boxes, box_index, pair_i, pair_j, pair_rank, pair_cap, category_mask & LIGHT_CATS, index_offset, hits,
                hit_cap, counters, wl_light, 4, 0);
        if (category_mask & FR3D_CAT_BACKBONE)
            fr3d::classify_pairs_tpp_kernel<FR3D_BPH_MINB, FR3D_CAT_BACKBONE, false><<<tpp_grid, 128, 0, st>>>(
                *nts, boxes, box_index, pair_i, pair_j, pair_rank, pair_cap, FR3D_CAT_BACKBONE, index_offset, hits,
                hit_cap, counters, wl_bph, 7, 0, screen64 ? nullptr : wl_bph_combos);
        // stacking and pairing lists: records staged in shared memory in two phases (classify_list2.cuh)
        constexpr int ts = fr3d::L2_WARPS_STACK * 32, tp = fr3d::L2_WARPS_PAIR * 32;
        const int grid_s = clampg((pair_cap + ts - 1) / ts, (long long)sm_count * (mult ? mult : 1));
        const int grid_p = clampg((pair_cap + tp - 1) / tp, (long long)sm_count * (mult ? mult : 1));
        if ((category_mask & FR3D_CAT_STACKING) && stack64) {
            fr3d::classify_list2_kernel<false, fr3d::L2_WARPS_STACK>
                <<<grid_s, ts, fr3d::l2_smem_bytes<false, fr3d::L2_WARPS_STACK>(), st>>>(
                    *nts, boxes, box_index, pair_i, pair_j, pair_rank, pair_cap, FR3D_CAT_STACKING, index_offset, hits,
                    hit_cap, counters, wl_stack, 5);
        } else if (category_mask & FR3D_CAT_STACKING) {
            // single precision where every comparison is certain, the undecided pairs through the fp64 kernel
            // (classify_stack32.cuh); the undecided list and its length live in the workspace
            int32_t *wl_unc = reinterpret_cast<int32_t *>(wsp + 4 * list_bytes);
            unsigned long long *unc_count = reinterpret_cast<unsigned long long *>(
                wsp + 6 * list_bytes + align_up(sizeof(float) * fr3d::S32_REC * (size_t)nts->n_nt, 256));
            CUDA_TRY(cudaMemsetAsync(unc_count, 0, sizeof(unsigned long long), st));
            const float *srec = reinterpret_cast<const float *>(wsp + 6 * list_bytes);
            constexpr int t32 = fr3d::ST32_WARPS * 32;
            fr3d::classify_stack32_kernel<<<clampg((pair_cap + t32 - 1) / t32, (long long)sm_count * (mult ? mult : 1)),
                                            t32, fr3d::ST32_SMEM, st>>>(
                srec, pair_i, pair_j, pair_rank, pair_cap, index_offset, hits, hit_cap, counters, wl_stack, 5, wl_unc,
                unc_count);
            // few pairs: the warp-per-pair kernel has by far the shortest latency per pair
            fr3d::classify_pairs_pm_kernel<<<sm_count * FR3D_CLASSIFY_MIN_BLOCKS, fr3d::WARPS_PER_BLOCK * 32, 0, st>>>(
                *nts, boxes, box_index, pair_i, pair_j, pair_rank, pair_cap, FR3D_CAT_STACKING | PM_CAT_NO_PAIRING,
                index_offset, hits, hit_cap, counters, wl_unc, reinterpret_cast<const int64_t *>(unc_count));
        }
        fr3d::classify_list2_kernel<true, fr3d::L2_WARPS_PAIR>
            <<<grid_p, tp, fr3d::l2_smem_bytes<true, fr3d::L2_WARPS_PAIR>(), st>>>(
                *nts, boxes, box_index, pair_i, pair_j, pair_rank, pair_cap, category_mask & FR3D_CAT_COPLANAR,
                index_offset, hits, hit_cap, counters, wl_pair, 6);
    } else if (use_mode == 1) {
        fr3d::classify_pairs_tpp_kernel<FR3D_TPP_MIN_BLOCKS, 0xFFFFFFFFu, true><<<clampg((pair_cap + 127) / 128, tpp_grid_max), 128, 0, st>>>(
            *nts, boxes, box_index, pair_i, pair_j, pair_rank, pair_cap, category_mask, index_offset, hits, hit_cap,
            counters, nullptr, 0, 1);
    } else {
        fr3d::classify_pairs_pm_kernel<<<clampg((pair_cap + per_block - 1) / per_block, pm_grid_max),
                                         fr3d::WARPS_PER_BLOCK * 32, 0, st>>>(
            *nts, boxes, box_index, pair_i, pair_j, pair_rank, pair_cap, category_mask, index_offset, hits, hit_cap,
            counters, nullptr);
    }
    CUDA_TRY(cudaGetLastError());
    return FR3D_OK;
}

namespace {
struct TailCarve {
    fr3d::TailWs ws;
    size_t bytes;
};
TailCarve carve_tail(void *workspace, int64_t hit_cap, int32_t n_nt, int32_t n_struct, int64_t n_ends) {
    TailCarve c;
    char *p = static_cast<char *>(workspace);
    size_t off = 0;
    auto take = [&](size_t bytes) { char *q = p ? p + off : nullptr; off += align_up(bytes ? bytes : 1, 256); return q; };
    const size_t H = (size_t)(hit_cap > 0 ? hit_cap : 0), n = (size_t)(n_nt > 0 ? n_nt : 0), S = (size_t)(n_struct > 0 ? n_struct : 0);
    fr3d::TailWs &w = c.ws;
    w.seg_cnt = reinterpret_cast<int32_t *>(take(4 * S));
    w.seg_off = reinterpret_cast<int32_t *>(take(4 * (S + 1)));
    w.seg_cur = reinterpret_cast<int32_t *>(take(4 * S));
    w.nt_struct = reinterpret_cast<int32_t *>(take(4 * n));
    w.tmpk = reinterpret_cast<unsigned long long *>(take(16 * H));
    w.tmpv = reinterpret_cast<uint32_t *>(take(8 * H));
    w.tmpc = reinterpret_cast<unsigned long long *>(take(8 * n));
    w.overflow = reinterpret_cast<int32_t *>(take(4));
    w.hkey = reinterpret_cast<unsigned long long *>(take(8 * H));
    w.hval = reinterpret_cast<uint32_t *>(take(4 * H));
    w.ekey = reinterpret_cast<unsigned long long *>(take(16 * H));
    w.eval = reinterpret_cast<uint32_t *>(take(8 * H));
    w.e_u = reinterpret_cast<int32_t *>(take(8 * H));
    w.e_v = reinterpret_cast<int32_t *>(take(8 * H));
    w.e_lab = reinterpret_cast<int32_t *>(take(8 * H));
    w.e_flag = reinterpret_cast<uint8_t *>(take(2 * H));
    w.akey = reinterpret_cast<unsigned long long *>(take(16 * H));
    w.aval = reinterpret_cast<uint32_t *>(take(8 * H));
    w.alab = reinterpret_cast<int32_t *>(take(8 * H));
    w.aa = reinterpret_cast<int32_t *>(take(8 * H));
    w.ab = reinterpret_cast<int32_t *>(take(8 * H));
    w.across = reinterpret_cast<int32_t *>(take(8 * H));
    w.first = reinterpret_cast<uint32_t *>(take(4 * n));
    w.gstart = reinterpret_cast<int32_t *>(take(4 * n));
    w.gcount = reinterpret_cast<int32_t *>(take(4 * n));
    w.ckey = reinterpret_cast<unsigned long long *>(take(8 * n));
    w.ends = reinterpret_cast<int32_t *>(take(4 * (size_t)(n_ends > 0 ? n_ends : 0)));
    c.bytes = off;
    return c;
}
}  // namespace

size_t fr3d_tail_workspace(int64_t hit_cap, int32_t n_nt, int32_t n_struct, int64_t n_ends) {
    return carve_tail(nullptr, hit_cap, n_nt, n_struct, n_ends).bytes;
}

int fr3d_annotation_tail(const fr3d_nts_t *nts, const fr3d_tail_ident_t *ident, const fr3d_tail_tables_t *tables,
                         const fr3d_hit_t *hits, int64_t hit_cap, const int64_t *counters, int32_t index_offset,
                         void *workspace, size_t workspace_bytes, fr3d_row_t *rows, int64_t row_cap,
                         int32_t *row_start, int32_t *row_count, int64_t *tail_counters, void *stream) {
    if (!nts || !ident || !tables || !hits || !counters || !workspace || !rows || !row_start || !row_count || !tail_counters)
        return fail(FR3D_E_ARG, "NULL argument");
    if (!ident->struct_off || !ident->struct_chain_off || !ident->chain_ends_off || !ident->nt_chain || !ident->nt_cov ||
        !tables->labels || !tables->boxes || !tables->slot_label)
        return fail(FR3D_E_ARG, "NULL table pointer");
    if (tables->n_labels <= 0 || tables->n_labels > fr3d::TAIL_MAX_LABELS) return fail(FR3D_E_ARG, "n_labels must be in [1, 1024]");
    if (hit_cap < 0 || row_cap < 0 || ident->n_struct < 0) return fail(FR3D_E_ARG, "negative size");
    if (2 * hit_cap >= (1ll << 31)) return fail(FR3D_E_ARG, "hit_cap must stay below 2^30 per call (split the batch)");
    int sm_count = 148;
    if (int rc = require_device(false, &sm_count)) return rc;
    if (ident->n_ends < 0 || ident->n_chains < 0) return fail(FR3D_E_ARG, "negative size");
    if (workspace_bytes < fr3d_tail_workspace(hit_cap, nts->n_nt, ident->n_struct, ident->n_ends))
        return fail(FR3D_E_ARG, "workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    fr3d::TailWs ws = carve_tail(workspace, hit_cap, nts->n_nt, ident->n_struct, ident->n_ends).ws;
    const int S = ident->n_struct;
    const int reset_threads = (S + 1 > nts->n_nt ? S + 1 : nts->n_nt);
    fr3d::tail_reset_kernel<<<(reset_threads + 255) / 256, 256, 0, st>>>(ws, *ident, nts->n_nt, hit_cap, tail_counters);
    if (S > 0) {
        const long long want = (hit_cap + 255) / 256;
        const long long cap_blocks = (long long)sm_count * 8;
        const int grid = (int)(want < 1 ? 1 : (want < cap_blocks ? want : cap_blocks));
        fr3d::tail_bin_kernel<<<grid, 256, 0, st>>>(hits, hit_cap, counters, index_offset, *ident, ws, tail_counters);
        fr3d::tail_scan_kernel<<<1, 1024, 0, st>>>(ws, S);                        // both return at once unless a structure
        fr3d::tail_scatter_kernel<<<grid, 256, 0, st>>>(hits, hit_cap, counters, index_offset, *ident, ws, tail_counters);
        // more (small) structures than 256-thread blocks fit at once: 128-thread blocks, all structures in flight together
        if (S > sm_count * FR3D_TAIL_MINB && nts->n_nt / S < fr3d::TAIL_BIG_NT) {
            const int blocks = S < sm_count * 8 ? S : sm_count * 8;
            fr3d::tail_structure_kernel<fr3d::TAIL_THREADS_MANY, 8, false>
                <<<blocks, fr3d::TAIL_THREADS_MANY, fr3d::tail_smem_bytes<fr3d::TAIL_THREADS_MANY>(), st>>>(
                    *nts, *ident, *tables, hits, index_offset, ws, rows, row_cap, row_start, row_count, tail_counters);
        } else {
            const int blocks = S < sm_count * FR3D_TAIL_MINB ? S : sm_count * FR3D_TAIL_MINB;
            fr3d::tail_structure_kernel<fr3d::TAIL_THREADS, FR3D_TAIL_MINB, false>
                <<<blocks, fr3d::TAIL_THREADS, fr3d::TAIL_SMEM, st>>>(*nts, *ident, *tables, hits, index_offset, ws, rows, row_cap,
                                                                      row_start, row_count, tail_counters);
        }
        if (ident->max_struct_nt >= fr3d::TAIL_BIG_NT) {                          // the large structures: 1024-thread blocks
            const int big_blocks = S < sm_count ? S : sm_count;
            fr3d::tail_structure_kernel<fr3d::TAIL_THREADS_BIG, 1, true>
                <<<big_blocks, fr3d::TAIL_THREADS_BIG, fr3d::tail_smem_bytes<fr3d::TAIL_THREADS_BIG>(), st>>>(
                    *nts, *ident, *tables, hits, index_offset, ws, rows, row_cap, row_start, row_count, tail_counters);
        }
    }
    CUDA_TRY(cudaGetLastError());
    return FR3D_OK;
}

size_t fr3d_radius_search_workspace(int32_t n_points) { return fr3d_pair_search_workspace(n_points); }

int fr3d_radius_search(const double *xyz, const int32_t *model, int32_t n_points, double max_cutoff, void *workspace,
                       size_t workspace_bytes, int32_t *pair_a, int32_t *pair_b, int32_t *pair_rank, double *dist2,
                       int64_t pair_cap, int64_t *counters, void *stream) {
    if (!xyz || !model || !workspace || !pair_a || !pair_b || !pair_rank || !dist2 || !counters)
        return fail(FR3D_E_ARG, "NULL argument");
    if (!(max_cutoff > 0)) return fail(FR3D_E_ARG, "max_cutoff must be positive");
    if (workspace_bytes < fr3d_radius_search_workspace(n_points)) return fail(FR3D_E_ARG, "workspace too small");
    if ((long long)n_points >= MAX_NT_INDEX) return fail(FR3D_E_ARG, "more than 2^27 points in one call (the list-order rank is int32)");
    int sm_count = 148;
    if (int rc = require_device(false, &sm_count)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaMemsetAsync(counters, 0, 8 * sizeof(int64_t), st));
    if (n_points <= 0) return FR3D_OK;
    fr3d::SearchWs ws = carve(workspace, n_points);
    const uint32_t T = ws.table_mask + 1;
    const int block = 256;
    fr3d::search_reset_kernel<<<(T + block - 1) / block, block, 0, st>>>(ws);
    fr3d::rs_assign_kernel<<<(n_points + block - 1) / block, block, 0, st>>>(xyz, model, n_points, max_cutoff, ws, counters);
    fr3d::cell_offsets_kernel<<<(T + fr3d::CELL_OFFSETS_THREADS - 1) / fr3d::CELL_OFFSETS_THREADS, fr3d::CELL_OFFSETS_THREADS, 0, st>>>(ws, counters);
    fr3d::rs_fill_kernel<<<(n_points + block - 1) / block, block, 0, st>>>(n_points, ws);
    // one warp per (occupied cell, stencil entry); the cell count lives on the device: size for the worst case,
    // capped at a few resident waves (the kernel strides over the tasks)
    const long long want = ((long long)n_points * 14 + fr3d::RS_WARPS - 1) / fr3d::RS_WARPS;
    const long long cap_blocks = (long long)sm_count * 32;
    fr3d::rs_pairs_kernel<<<(int)(want < cap_blocks ? want : cap_blocks), fr3d::RS_WARPS * 32, 0, st>>>(
        xyz, model, ws, max_cutoff * max_cutoff, pair_a, pair_b, pair_rank, dist2, pair_cap, counters);
    CUDA_TRY(cudaGetLastError());
    return FR3D_OK;
}

int fr3d_bond_orientation(const fr3d_nts_t *nts, const int32_t *host_ring_slot, double *chi_degree, int32_t *orientation,
                          void *stream) {
    if (!nts || !host_ring_slot || !chi_degree || !orientation) return fail(FR3D_E_ARG, "NULL argument");
    if (int rc = require_device(false, nullptr)) return rc;
    if (nts->n_nt <= 0) return FR3D_OK;
    fr3d::RingSlots ring;
    for (int p = 0; p < FR3D_N_PARENTS; ++p) {
        if (host_ring_slot[p] < 0 || host_ring_slot[p] >= FR3D_BASE_SLOTS) return fail(FR3D_E_ARG, "ring slot out of range");
        ring.s[p] = host_ring_slot[p];
    }
    const int block = 128;
    fr3d::bond_orientation_kernel<<<(nts->n_nt + block - 1) / block, block, 0, static_cast<cudaStream_t>(stream)>>>(
        *nts, ring, chi_degree, orientation);
    CUDA_TRY(cudaGetLastError());
    return FR3D_OK;
}

int fr3d_kabsch_batched(const double *set1, const double *set2, const double *w, const int32_t *off, int32_t n_prob,
                        double *U, double *mean1, double *mean2, double *rmsd, double *sse, double *new1,
                        void *stream) {
    if (!set1 || !set2 || !off) return fail(FR3D_E_ARG, "NULL argument");      // every output may be NULL
    if (int rc = require_device(false, nullptr)) return rc;
    if (n_prob <= 0) return FR3D_OK;
    const int block = fr3d::KABSCH_WARPS * 32;
    const size_t smem = w ? fr3d::KABSCH_SMEM : fr3d::KABSCH_SMEM / 7 * 6;          // no weights: no weight rows staged
    fr3d::kabsch_batched_kernel<<<(n_prob + block - 1) / block, block, smem, static_cast<cudaStream_t>(stream)>>>(
        set1, set2, w, off, n_prob, U, mean1, mean2, rmsd, sse, new1);
    CUDA_TRY(cudaGetLastError());
    return FR3D_OK;
}

static int launch_all_vs_all(const double *centers, const double *rots, const uint8_t *has_rot, int32_t N, int32_t n,
                             double angle_weight, int32_t row0, int32_t row1, int32_t col0, int32_t ld, double *out,
                             void *stream) {
    if (!centers || !rots || !out) return fail(FR3D_E_ARG, "NULL argument");
    if (n < 2 || n > fr3d::DISC_MAXN) return fail(FR3D_E_ARG, "n must be in [2, 16]");
    if (row0 < 0 || row1 > N || row0 > row1) return fail(FR3D_E_ARG, "bad row range");
    if (col0 < 0 || col0 % fr3d::DISC_TILE) return fail(FR3D_E_ARG, "the first column must be a multiple of 32");
    if (int rc = require_device(false, nullptr)) return rc;
    if (row0 == row1 || N == 0 || col0 >= N) return FR3D_OK;
    const size_t smem = fr3d::disc_smem_bytes(n);        // opted in for n = DISC_MAXN by fr3d_init
    dim3 block(32, 8);
    dim3 grid((N - col0 + fr3d::DISC_TILE - 1) / fr3d::DISC_TILE, (row1 - row0 + fr3d::DISC_TILE - 1) / fr3d::DISC_TILE);
    fr3d::discrepancy_all_vs_all_kernel<<<grid, block, smem, static_cast<cudaStream_t>(stream)>>>(
        centers, rots, has_rot, N, n, angle_weight, row0, row1, col0, ld, out);
    CUDA_TRY(cudaGetLastError());
    return FR3D_OK;
}

int fr3d_discrepancy_all_vs_all(const double *centers, const double *rots, const uint8_t *has_rot, int32_t N,
                                int32_t n, double angle_weight, int32_t row0, int32_t row1, double *out,
                                void *stream) {
    return launch_all_vs_all(centers, rots, has_rot, N, n, angle_weight, row0, row1, 0, N, out, stream);
}

int fr3d_discrepancy_rows_upper(const double *centers, const double *rots, const uint8_t *has_rot, int32_t N,
                                int32_t n, double angle_weight, int32_t row0, int32_t row1, double *out,
                                void *stream) {
    if (row0 % fr3d::DISC_TILE) return fail(FR3D_E_ARG, "row0 must be a multiple of 32");
    return launch_all_vs_all(centers, rots, has_rot, N, n, angle_weight, row0, row1, row0, N - row0, out, stream);
}

int fr3d_discrepancy_assemble(const double *block_upper, int32_t N, int32_t row0, int32_t row1, double *matrix,
                              void *stream) {
    if (!block_upper || !matrix) return fail(FR3D_E_ARG, "NULL argument");
    if (row0 < 0 || row1 > N || row0 > row1 || row0 % fr3d::DISC_TILE) return fail(FR3D_E_ARG, "bad row range");
    if (int rc = require_device(false, nullptr)) return rc;
    if (row0 == row1) return FR3D_OK;
    dim3 block(32, 8);
    dim3 grid((N - row0 + fr3d::DISC_TILE - 1) / fr3d::DISC_TILE, (row1 - row0 + fr3d::DISC_TILE - 1) / fr3d::DISC_TILE);
    fr3d::discrepancy_assemble_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(block_upper, N, row0, row1, matrix);
    CUDA_TRY(cudaGetLastError());
    return FR3D_OK;
}

int fr3d_fp64_fma_probe(int64_t iters, double *sink, int64_t *flops, void *stream) {
    if (!sink || !flops || iters <= 0) return fail(FR3D_E_ARG, "bad argument");
    int sm_count = 148;
    if (int rc = require_device(false, &sm_count)) return rc;
    const int blocks = sm_count * 8;
    fr3d::fp64_fma_probe_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(iters, 1.0, sink);
    CUDA_TRY(cudaGetLastError());
    *flops = (int64_t)blocks * 256 * iters * 16;
    return FR3D_OK;
}

int fr3d_discrepancy_one_vs_many(const double *q_centers, const double *q_rots, const uint8_t *q_has_rot,
                                 const double *centers, const double *rots, const uint8_t *has_rot, int32_t M,
                                 int32_t n, double cutoff, double angle_weight, const double *center_weight,
                                 double *out, void *stream) {
    if (!q_centers || !q_rots || !centers || !rots || !out) return fail(FR3D_E_ARG, "NULL argument");
    if (n < 2 || n > fr3d::DISC_MAXN) return fail(FR3D_E_ARG, "n must be in [2, 16]");
    if (int rc = require_device(false, nullptr)) return rc;
    if (M <= 0) return FR3D_OK;
    const int block = 128;
    fr3d::discrepancy_one_vs_many_kernel<<<(M + block - 1) / block, block, 0, static_cast<cudaStream_t>(stream)>>>(
        q_centers, q_rots, q_has_rot, centers, rots, has_rot, M, n, cutoff, angle_weight, center_weight, out);
    CUDA_TRY(cudaGetLastError());
    return FR3D_OK;
}

// ---- K8: ordering by similarity (csrc/ordering.cuh) ----------------------------------------------------------------
namespace {
struct OrderCarve { fr3d::LinkWs link; fr3d::PathWs path; size_t bytes; };
OrderCarve carve_order(void *workspace, int32_t n, int32_t reps) {
    OrderCarve c;
    size_t off = 0;
    char *base = static_cast<char *>(workspace);
    auto take = [&](size_t bytes) { char *p = base ? base + off : nullptr; off += align_up(bytes, 256); return p; };
    const size_t N = (size_t)(n > 0 ? n : 0), R = (size_t)(reps > 0 ? reps : 0);
    c.link.W = reinterpret_cast<double *>(take(8 * N * N));
    c.link.m_h = reinterpret_cast<double *>(take(8 * N));
    int32_t **ints[] = {&c.link.size, &c.link.chain, &c.link.next, &c.link.head, &c.link.tail, &c.link.m_head,
                        &c.link.m_nx, &c.link.m_ny, &c.link.leaf, &c.link.pos};
    for (int32_t **p : ints) *p = reinterpret_cast<int32_t *>(take(4 * N));
    c.path.edge = reinterpret_cast<double *>(take(8 * 2 * N * R));
    c.path.path = reinterpret_cast<int32_t *>(take(4 * 2 * N * R));
    c.bytes = off;
    return c;
}
}  // namespace

size_t fr3d_order_workspace(int32_t n, int32_t reps) { return carve_order(nullptr, n, reps).bytes; }

int fr3d_tree_penalty(const double *distance, int32_t n, double *penalty, int32_t *flags, void *workspace,
                      size_t workspace_bytes, void *stream) {
    if (!distance || !penalty || !flags || !workspace) return fail(FR3D_E_ARG, "NULL argument");
    if (n < 0 || n > 46340) return fail(FR3D_E_ARG, "n must be in [0, 46340]");
    if (workspace_bytes < fr3d_order_workspace(n, 0)) return fail(FR3D_E_ARG, "workspace too small");
    int sm_count = 148;
    if (int rc = require_device(false, &sm_count)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaMemsetAsync(flags, 0, sizeof(int32_t), st));
    if (n == 0) return FR3D_OK;
    CUDA_TRY(cudaMemsetAsync(penalty, 0, sizeof(double) * (size_t)n * n, st));
    if (n == 1) return FR3D_OK;
    fr3d::LinkWs ws = carve_order(workspace, n, 0).link;
    const long long total = (long long)n * n;
    const long long want = (total + 255) / 256;
    fr3d::linkage_prepare_kernel<<<(int)(want < sm_count * 8 ? want : sm_count * 8), 256, 0, st>>>(distance, n, ws.W, flags);
    fr3d::linkage_chain_kernel<<<1, fr3d::ORD_THREADS, 0, st>>>(ws, n, flags);
    const int slices = n >= 2048 ? 64 : (n >= 256 ? 8 : 1);
    fr3d::cophenetic_fill_kernel<<<dim3(n - 1, slices), 256, 0, st>>>(ws, n, penalty, flags);
    CUDA_TRY(cudaGetLastError());
    return FR3D_OK;
}

int fr3d_penalized_distance(const double *distance, const double *penalty, int32_t n, int32_t mode, double strength,
                            double *sums, double *out, void *stream) {
    if (!distance || !penalty || !sums || !out) return fail(FR3D_E_ARG, "NULL argument");
    if (n < 0 || n > 46340 || mode < 0 || mode > 2) return fail(FR3D_E_ARG, "bad n or mode");
    int sm_count = 148;
    if (int rc = require_device(false, &sm_count)) return rc;
    if (n == 0) return FR3D_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long total = (long long)n * n;
    if (mode >= 1) fr3d::sequential_sum_kernel<<<2, 256, 0, st>>>(distance, penalty, total, sums);
    const long long want = (total + 255) / 256;
    fr3d::penalized_distance_kernel<<<(int)(want < sm_count * 8 ? want : sm_count * 8), 256, 0, st>>>(
        distance, penalty, total, mode, strength, sums, out);
    CUDA_TRY(cudaGetLastError());
    return FR3D_OK;
}

int fr3d_order_paths(const double *distance, int32_t n, const int32_t *start_orders, int32_t reps, int32_t mode,
                     int32_t *orders, double *scores, double *reductions, void *workspace, size_t workspace_bytes,
                     void *stream) {
    if (!distance || !start_orders || !orders || !scores || !reductions || !workspace) return fail(FR3D_E_ARG, "NULL argument");
    if (n < 2 || n > 46340) return fail(FR3D_E_ARG, "n must be in [2, 46340]");
    if (reps < 1 || mode < 1 || mode > 3) return fail(FR3D_E_ARG, "reps >= 1 and mode in {1 greedy, 2 two-opt, 3 both}");
    if (workspace_bytes < fr3d_order_workspace(n, reps)) return fail(FR3D_E_ARG, "workspace too small");
    if (int rc = require_device(false, nullptr)) return rc;
    fr3d::PathWs ws = carve_order(workspace, n, reps).path;
    fr3d::order_paths_kernel<<<reps, fr3d::ORD_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        distance, n, start_orders, mode, ws, orders, scores, reductions);
    CUDA_TRY(cudaGetLastError());
    return FR3D_OK;
}

int fr3d_reorder_symmetric(const double *distance, int32_t n, const int32_t *order, int32_t zero_diagonal, double *out,
                           void *stream) {
    if (!distance || !order || !out) return fail(FR3D_E_ARG, "NULL argument");
    if (n < 0 || n > 46340) return fail(FR3D_E_ARG, "n must be in [0, 46340]");
    int sm_count = 148;
    if (int rc = require_device(false, &sm_count)) return rc;
    if (n == 0) return FR3D_OK;
    const long long want = ((long long)n * n + 255) / 256;
    fr3d::reorder_symmetric_kernel<<<(int)(want < sm_count * 8 ? want : sm_count * 8), 256, 0,
                                     static_cast<cudaStream_t>(stream)>>>(distance, n, order, zero_diagonal, out);
    CUDA_TRY(cudaGetLastError());
    return FR3D_OK;
}

// ---- K9: FR3D search, constraint intersection (csrc/possibilities.cuh) ---------------------------------------------
int fr3d_possibilities_first(const fr3d_search_query_t *dev_query, const int32_t *pairs, int64_t n_pairs, int32_t emit,
                             int64_t *counts, int32_t *out_frags, double *out_errs, int64_t out_cap, void *stream) {
    if (!dev_query || !counts || (n_pairs > 0 && !pairs)) return fail(FR3D_E_ARG, "NULL argument");
    if (n_pairs < 0 || (emit && (!out_frags || out_cap < 0))) return fail(FR3D_E_ARG, "bad size or NULL output");
    int sm_count = 148;
    if (int rc = require_device(false, &sm_count)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long want = (n_pairs + 255) / 256;
    const int grid = (int)(want < 1 ? 1 : (want < sm_count * 8 ? want : sm_count * 8));
    static_assert(sizeof(long long) == sizeof(int64_t), "int64");
    long long *cnt = reinterpret_cast<long long *>(counts);
    if (!emit) {
        fr3d::poss_first_pairs_kernel<<<grid, 256, 0, st>>>(dev_query, pairs, n_pairs, cnt);
        fr3d::poss_scan_kernel<<<1, 1024, 0, st>>>(cnt, n_pairs);
    } else {
        fr3d::poss_first_emit_kernel<<<grid, 256, 0, st>>>(dev_query, pairs, n_pairs, cnt, out_frags, out_errs, out_cap);
    }
    CUDA_TRY(cudaGetLastError());
    return FR3D_OK;
}

int fr3d_possibilities_extend(const fr3d_search_query_t *dev_query, int32_t level, const int32_t *frags, const double *errs,
                              int64_t n_frags, int32_t emit, int64_t *counts, int32_t *out_frags, double *out_errs,
                              int64_t out_cap, int32_t *flags, void *stream) {
    if (!dev_query || !counts || !flags || (n_frags > 0 && !frags)) return fail(FR3D_E_ARG, "NULL argument");
    if (level < 1 || level >= FR3D_SEARCH_MAX_POSITIONS - 1) return fail(FR3D_E_ARG, "level must be in [1, 14]");
    if (n_frags < 0 || (emit && (!out_frags || out_cap < 0))) return fail(FR3D_E_ARG, "bad size or NULL output");
    int sm_count = 148;
    if (int rc = require_device(false, &sm_count)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long want = (n_frags + fr3d::POSS_WARPS - 1) / fr3d::POSS_WARPS;
    const int grid = (int)(want < 1 ? 1 : (want < sm_count * 16 ? want : sm_count * 16));
    long long *cnt = reinterpret_cast<long long *>(counts);
    if (!emit) {
        fr3d::poss_extend_kernel<false><<<grid, fr3d::POSS_WARPS * 32, 0, st>>>(dev_query, level, frags, errs, n_frags, cnt,
                                                                                nullptr, nullptr, 0, flags);
        fr3d::poss_scan_kernel<<<1, 1024, 0, st>>>(cnt, n_frags);
    } else {
        fr3d::poss_extend_kernel<true><<<grid, fr3d::POSS_WARPS * 32, 0, st>>>(dev_query, level, frags, errs, n_frags, cnt,
                                                                               out_frags, out_errs, out_cap, flags);
    }
    CUDA_TRY(cudaGetLastError());
    return FR3D_OK;
}

}  // extern "C"
