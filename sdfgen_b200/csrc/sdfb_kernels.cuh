// sdfb_kernels.cuh -- launch wrappers shared between the kernel TUs and the C-ABI host layer.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "sdfb_math.cuh"

namespace sdfb {

// Geometry of the slab a plan owns.  Planes of `cells` are indexed p = k - k_lo + 1, so plane 0 and
// plane nkl+1 are the halo planes (copies of the neighbouring slabs' boundary planes, or the
// initial cell when the slab touches the grid boundary).
struct Grid {
    int ni, nj, nk;        // global extent
    int k_lo, k_hi;        // owned planes [k_lo, k_hi)
    float dx, ox, oy, oz;
    int band;
    __host__ __device__ int nkl() const { return k_hi - k_lo; }
    __host__ __device__ int64_t plane() const { return (int64_t)ni * nj; }
    __host__ __device__ int64_t slab_voxels() const { return plane() * nkl(); }
    __host__ __device__ int64_t cell_count() const { return plane() * (nkl() + 2); }
    // index into cells of global voxel (i,j,k), k in [k_lo-1, k_hi]
    __host__ __device__ int64_t cidx(int i, int j, int k) const { return (int64_t)i + (int64_t)ni * ((int64_t)j + (int64_t)nj * (int64_t)(k - k_lo + 1)); }
    // index into slab-local dense arrays (counts, phi out), k in [k_lo, k_hi)
    __host__ __device__ int64_t vidx(int i, int j, int k) const { return (int64_t)i + (int64_t)ni * ((int64_t)j + (int64_t)nj * (int64_t)(k - k_lo)); }
};

// Integer extents of one triangle on the grid (sdfb_band.cu): exact-band box and yz lattice range of the x-ray test,
// both inclusive and clipped to the slab's planes.
struct __align__(16) TriExt { int i0, i1, j0, j1, k0, k1, cj0, cj1, ck0, ck1, pad0, pad1; };

// Development knobs, read from the environment ONCE per plan (sdfb_plan_create) -- none is needed in production.
struct Tuning {
    int relax_from = -1;          // SDFB_RELAX_FROM: first sweep index run by the relaxation schedule (-1: default 8)
    int fuse_pass = -1;           // SDFB_FUSE_PASS: 0 one launch per first-pass sweep, 1 fuse launches of >= 300 M voxels too
    int minb = 0;                 // SDFB_MINB: 3 / 4 = register bound (CTAs per SM) of the column kernels, 0 = by launch size
    int max_occ = 0;              // SDFB_MAX_OCC: cap on resident column CTAs per SM (experiments)
    int cta_queue = -1;           // SDFB_CTA_QUEUE: 0/1 force the warp-private / column-wide evaluation queue
    int cta_queue_until = 8;      // SDFB_CTA_QUEUE_UNTIL: first sweep that uses warp-private queues (default: the second pass)
    int relax_list_cap = 0;       // SDFB_RELAX_LIST_CAP: work-list capacity (tests force the bitmap fallback)
    long long relax_heavy_limit = -1;   // SDFB_RELAX_HEAVY_LIMIT: work-list entries before a sweep is handed back to the columns
    int relax_scan_from = 13;     // SDFB_RELAX_SCAN_FROM: first sweep whose round 0 uses the lean scan kernel
    int relax_debug = 0;          // SDFB_RELAX_DEBUG: per-sweep round statistics on stderr
    int lookahead = 1;            // SDFB_LOOKAHEAD: 0 = every relaxation sweep scans the grid for itself (no lookahead window)
    int look_cap = 0;             // SDFB_LOOK_CAP: capacity of the window's lists (tests force the overflow -> dense round 0 path)
    int early_copy = 1;           // SDFB_EARLY_COPY: 0 = the one-shot call downloads phi after the last sweep, not during the second pass
    int col_shape = 0;            // SDFB_COL_SHAPE: 12 / 16 = force the 8 x 12 / 8 x 16 build of the column schedule (0: by launch size)
    int order_w = -1;             // SDFB_ORDER_W: ticket order of fused launches by the key w*J + K (1 = anti-diagonals, >= NK = row by row)
    int link_timeout_s = 20;      // SDFB_LINK_TIMEOUT_S: watchdog of the cross-GPU waits (the kernel traps instead of hanging)
    int link_debug = 0;           // SDFB_LINK_DEBUG: TIMING EXPERIMENTS ONLY, results are wrong -- 1: boundary cells are stored into a
                                  // local dummy plane instead of the neighbour's memory, 2: device-scope fence before the link flag,
                                  // 4: do not wait for the upstream neighbour's flags
    int link_trace = 0;           // SDFB_LINK_TRACE: record first-column-start / last-column-end times of every sweep (sdfb_plan_link_trace)
};
Tuning tuning_from_env();

// Exact multi-GPU mode (linked k-slabs): per-sweep hand-over of the slab boundary plane between neighbouring plans.
// A sweep reads only offsets 0 and -dk in k (cpu_lib/makelevelset3.cpp:143-149), so the slab downstream of this one
// consumes this slab's last plane column by column AS THE COLUMNS COMPLETE: the columns of the last K block store their
// boundary-plane cells into the neighbour's inbound buffer for this sweep (peer memory over NVLink) and publish their
// step count there with a system-scope fence; the neighbour's first K block polls those words exactly like it polls a
// column below it on the same GPU.  One inbound plane and one row of flags PER SWEEP INDEX (16 of each): a buffer is
// written by exactly one sweep of one run, so a neighbour that runs ahead (the next sweep with the same k direction)
// can never overwrite cells that are still being read.
constexpr int LINK_SWEEPS = 16;
struct LinkSweep {
    const uint64_t *halo_src = nullptr;             // [nj][ni] cells of the upstream slab's boundary plane (local memory, written by the peer)
    const unsigned long long *flag_src = nullptr;   // [NJ] run << 32 | steps completed by the upstream column (J, last K block)
    uint64_t *halo_dst = nullptr;                   // the downstream peer's halo_src of this sweep
    unsigned long long *flag_dst = nullptr;         // the downstream peer's flag_src of this sweep
};
struct LinkState {          // host side, per plan
    bool active = false;
    uint64_t *in_halo = nullptr;                    // LINK_SWEEPS planes, cudaMalloc (IPC-exportable)
    unsigned long long *in_flags = nullptr;         // LINK_SWEEPS x NJ words, same allocation
    uint64_t *peer_halo[2] = {nullptr, nullptr};    // [0] the slab below (k_lo - 1), [1] the slab above (k_hi)
    unsigned long long *peer_flags[2] = {nullptr, nullptr};
    void *peer_base[2] = {nullptr, nullptr};        // what cudaIpcOpenMemHandle returned (nullptr for same-process links)
    unsigned long long run = 0;                     // band() calls so far: flag words are run << 32 | steps
    int NJ = 0;
    size_t bytes = 0;
    unsigned long long *trace = nullptr;            // SDFB_LINK_TRACE: 2 x LINK_SWEEPS words (~start, end in globaltimer ns), else nullptr
};

struct Launches { uint64_t n = 0; };

// all launchers enqueue on `st` and return the number of kernels launched
int launch_init(uint64_t *cells, int64_t ncells, float init_phi, cudaStream_t st);
int launch_tri_prep(const uint32_t *tri, const float *xyz, uint64_t ntri, uint64_t nvert, TriRec *rec,
                    unsigned long long *bad, cudaStream_t st);
int launch_band(const TriRec *rec, uint64_t ntri, const Grid &g, uint32_t *units, TriExt *ext, uint64_t *prefix,
                uint64_t *block_sums, uint64_t *cells, int32_t *counts, float init_phi, cudaStream_t st);
int launch_sweep_levels(uint64_t *cells, const TriRec *rec, const Grid &g, int sweep_index,
                        unsigned long long *changed, cudaStream_t st);
int launch_sweep_columns(uint64_t *cells, const TriRec *rec, const Grid &g, int sweep_index,
                         unsigned long long *changed, uint32_t *progress, uint32_t epoch, cudaStream_t st,
                         const Tuning &tun, const unsigned int *run_if = nullptr, int max_ctas = 0);
// link: nullptr, or the plan's link state (then the launch hands its boundary plane over / waits for the neighbour's)
int launch_sweep_columns_fused(uint64_t *cells, const TriRec *rec, const Grid &g, int first, int count,
                               unsigned long long *changed, uint32_t *progress, size_t progress_words, uint32_t *epoch,
                               cudaStream_t st, const Tuning &tun, int max_ctas, const LinkState *link = nullptr);
size_t link_flag_words_per_sweep(const Grid &g);     // NJ of the column schedule
// the same schedule built with 8 x 12 columns (sdfb_sweep_columns_ek12.cu): four 160-thread CTAs per SM, for launches below
// 300 M voxels on unlinked plans
int launch_sweep_columns_ek12(uint64_t *cells, const TriRec *rec, const Grid &g, int sweep_index,
                              unsigned long long *changed, uint32_t *progress, uint32_t epoch, cudaStream_t st,
                              const Tuning &tun, const unsigned int *run_if = nullptr, int max_ctas = 0);
int launch_sweep_columns_fused_ek12(uint64_t *cells, const TriRec *rec, const Grid &g, int first, int count,
                                    unsigned long long *changed, uint32_t *progress, size_t progress_words, uint32_t *epoch,
                                    cudaStream_t st, const Tuning &tun, int max_ctas, const LinkState *link = nullptr);
size_t sweep_columns_progress_words_ek12(const Grid &g);
int launch_sign(const uint64_t *cells, const int32_t *counts, const Grid &g, bool apply_sign,
                bool kfastest, float *phi_out, cudaStream_t st);
int launch_unpack_tri(const uint64_t *cells, const Grid &g, bool kfastest, int32_t *tri_out, cudaStream_t st);
int launch_halo_refresh(uint64_t *cells, const Grid &g, cudaStream_t st);
int launch_verify_cells(const uint64_t *cells, const TriRec *rec, const Grid &g, float init_phi, unsigned long long *out, cudaStream_t st);
int launch_count_negative(const float *v, int64_t n, unsigned long long *out, cudaStream_t st);
int launch_relayout_i32(const int32_t *src, const Grid &g, int32_t *dst_kfastest, cudaStream_t st);

bool sweep_relax_supported(const Grid &g);
size_t sweep_relax_scratch_bytes(const Grid &g);
const unsigned int *sweep_relax_fallback_flag(const void *scratch);
// look: the sweep belongs to a lookahead window opened by launch_look_scan (its round 0 starts from the window's lists)
int launch_sweep_relax(uint64_t *cells, const TriRec *rec, const Grid &g, int sweep_index,
                       unsigned long long *changed, void *scratch, cudaStream_t st, const Tuning &tun, int max_ctas = 0,
                       bool look = false);
// one pass over the cells that finds, for each of the sweeps s_lo .. s_hi-1 (<= 8), the voxels a candidate can still
// improve; returns the number of launches (0: not applicable to this grid)
int launch_look_scan(const uint64_t *cells, const TriRec *rec, const Grid &g, int s_lo, int s_hi,
                     unsigned long long *changed, void *scratch, cudaStream_t st, const Tuning &tun, int max_ctas = 0);

// what the window changed, as {output index, value} patches for an output made from the cells before the window
int launch_look_patches(const uint64_t *cells, const float *phi_early, const Grid &g, bool kfastest, void *scratch, const Tuning &tun,
                        uint32_t *patch_idx, float *patch_val, uint32_t cap, unsigned int *head, cudaStream_t st);

size_t sweep_columns_progress_words(const Grid &g);

}  // namespace sdfb
