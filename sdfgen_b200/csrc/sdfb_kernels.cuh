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
                         const unsigned int *run_if = nullptr, int max_ctas = 0);
int launch_sweep_columns_fused(uint64_t *cells, const TriRec *rec, const Grid &g, int first, int count,
                               unsigned long long *changed, uint32_t *progress, size_t progress_words, uint32_t *epoch,
                               cudaStream_t st, int max_ctas);
int launch_sign(const uint64_t *cells, const int32_t *counts, const Grid &g, bool apply_sign,
                bool kfastest, float *phi_out, cudaStream_t st);
int launch_unpack_tri(const uint64_t *cells, const Grid &g, bool kfastest, int32_t *tri_out, cudaStream_t st);
int launch_halo_refresh(uint64_t *cells, const Grid &g, cudaStream_t st);
int launch_count_negative(const float *v, int64_t n, unsigned long long *out, cudaStream_t st);
int launch_relayout_i32(const int32_t *src, const Grid &g, int32_t *dst_kfastest, cudaStream_t st);

int launch_sweep_strips(uint64_t *cells, const TriRec *rec, const Grid &g, int sweep_index,
                        unsigned long long *changed, uint32_t *progress, uint32_t epoch, cudaStream_t st);

bool sweep_relax_supported(const Grid &g);
size_t sweep_relax_scratch_bytes(const Grid &g);
const unsigned int *sweep_relax_fallback_flag(const void *scratch);
int launch_sweep_relax(uint64_t *cells, const TriRec *rec, const Grid &g, int sweep_index,
                       unsigned long long *changed, void *scratch, cudaStream_t st, int max_ctas = 0);

size_t sweep_columns_progress_words(const Grid &g);
size_t sweep_strips_progress_words(const Grid &g);

}  // namespace sdfb
