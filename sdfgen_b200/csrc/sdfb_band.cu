// sdfb_band.cu -- phase A of make_level_set3 on sm_100a: grid init, per-triangle records, exact
// band distances and x-ray crossing counts (reference: cpu_lib/makelevelset3.cpp:196-236).
//
// The reference walks triangles serially and keeps, per voxel, the first strictly smaller distance.
// That equals the lexicographic minimum of (distance bits, triangle index), which is order
// independent, so every (triangle, voxel) pair is evaluated in parallel and resolved with one
// 64-bit atomicMin on the packed cell (SASS REDG.E.MIN.64).  Crossing counts are integer adds and
// commute, so they use atomicAdd.  Work is cut into warp-sized units of <= UNIT voxels / lattice
// points so that one huge triangle cannot serialise the phase: a prefix sum over per-triangle unit
// counts maps a unit number back to (triangle, offset) with a binary search.
#include <cstring>
#include "sdfb_kernels.cuh"

namespace sdfb {

namespace {

constexpr int UNIT = 256;          // voxels (or yz lattice points) per warp work unit
constexpr int SCAN_TILE = 2048;    // elements per block in the prefix sum

__global__ void k_init_cells(uint64_t *cells, int64_t n, uint64_t init_cell)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // two cells per 16-byte store, four stores in flight per thread
    ulonglong2 v = make_ulonglong2(init_cell, init_cell);
    int64_t n2 = n >> 1;
    int64_t c = i;
    for (; c + 3 * stride < n2; c += 4 * stride) {
        reinterpret_cast<ulonglong2 *>(cells)[c] = v;
        reinterpret_cast<ulonglong2 *>(cells)[c + stride] = v;
        reinterpret_cast<ulonglong2 *>(cells)[c + 2 * stride] = v;
        reinterpret_cast<ulonglong2 *>(cells)[c + 3 * stride] = v;
    }
    for (; c < n2; c += stride) reinterpret_cast<ulonglong2 *>(cells)[c] = v;
    if (i == 0 && (n & 1)) cells[n - 1] = init_cell;
}

// gather the three vertices of each triangle into its 48-byte record
// A vertex index >= nvert is undefined behaviour in the reference (it indexes x[] unchecked,
// cpu_lib/makelevelset3.cpp:205); here it must not become an illegal address that kills the CUDA context: the
// lowest offending triangle id is recorded in *bad (atomicMin, initialised to ~0) and the index is replaced by 0;
// the blocking calls that hand results to the host report SDFB_ERR_INVALID.
__global__ void k_tri_prep(const uint32_t *__restrict__ tri, const float *__restrict__ xyz, uint64_t ntri, uint64_t nvert,
                           TriRec *__restrict__ rec, unsigned long long *__restrict__ bad)
{
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntri) return;
    uint32_t p = tri[3 * t], q = tri[3 * t + 1], r = tri[3 * t + 2];
    if (p >= nvert || q >= nvert || r >= nvert) {
        atomicMin(bad, (unsigned long long)t);
        if (p >= nvert) p = 0;
        if (q >= nvert) q = 0;
        if (r >= nvert) r = 0;
    }
    rec[t] = make_tri_rec(F3{xyz[3 * (size_t)p], xyz[3 * (size_t)p + 1], xyz[3 * (size_t)p + 2]},
                          F3{xyz[3 * (size_t)q], xyz[3 * (size_t)q + 1], xyz[3 * (size_t)q + 2]},
                          F3{xyz[3 * (size_t)r], xyz[3 * (size_t)r + 1], xyz[3 * (size_t)r + 2]});
}

// Integer extents of one triangle on the grid: the exact-band box (cpu_lib/makelevelset3.cpp:206-212)
// and the yz lattice range of the x-ray test (:222-225), both clipped to the slab's planes.
struct TriBoxes {
    double fip, fjp, fkp, fiq, fjq, fkq, fir, fjr, fkr;
    int i0, i1, j0, j1, k0, k1;     // band box, inclusive
    int cj0, cj1, ck0, ck1;         // crossing lattice range, inclusive
    __device__ uint64_t band_voxels() const { return (k1 < k0) ? 0 : (uint64_t)(i1 - i0 + 1) * (uint64_t)(j1 - j0 + 1) * (uint64_t)(k1 - k0 + 1); }
    __device__ uint64_t cross_points() const { return (ck1 < ck0 || cj1 < cj0) ? 0 : (uint64_t)(cj1 - cj0 + 1) * (uint64_t)(ck1 - ck0 + 1); }
};

__device__ __forceinline__ double grid_coord(float x, float o, float dx)
{
    return __ddiv_rn(__dsub_rn((double)x, (double)o), (double)dx);
}
__device__ __forceinline__ double min3(double a, double b, double c) { return min_std(a, min_std(b, c)); }
__device__ __forceinline__ double max3(double a, double b, double c) { return max_std(a, max_std(b, c)); }

__device__ __forceinline__ void tri_boxes(const TriRec &t, const Grid &g, TriBoxes &b)
{
    b.fip = grid_coord(t.p.x, g.ox, g.dx); b.fjp = grid_coord(t.p.y, g.oy, g.dx); b.fkp = grid_coord(t.p.z, g.oz, g.dx);
    b.fiq = grid_coord(t.q.x, g.ox, g.dx); b.fjq = grid_coord(t.q.y, g.oy, g.dx); b.fkq = grid_coord(t.q.z, g.oz, g.dx);
    b.fir = grid_coord(t.r.x, g.ox, g.dx); b.fjr = grid_coord(t.r.y, g.oy, g.dx); b.fkr = grid_coord(t.r.z, g.oz, g.dx);
    double ilo = min3(b.fip, b.fiq, b.fir), ihi = max3(b.fip, b.fiq, b.fir);
    double jlo = min3(b.fjp, b.fjq, b.fjr), jhi = max3(b.fjp, b.fjq, b.fjr);
    double klo = min3(b.fkp, b.fkq, b.fkr), khi = max3(b.fkp, b.fkq, b.fkr);
    b.i0 = iclamp(wrap_add(d2i_trunc(ilo), -g.band), 0, g.ni - 1); b.i1 = iclamp(wrap_add(d2i_trunc(ihi), g.band + 1), 0, g.ni - 1);
    b.j0 = iclamp(wrap_add(d2i_trunc(jlo), -g.band), 0, g.nj - 1); b.j1 = iclamp(wrap_add(d2i_trunc(jhi), g.band + 1), 0, g.nj - 1);
    b.k0 = iclamp(wrap_add(d2i_trunc(klo), -g.band), 0, g.nk - 1); b.k1 = iclamp(wrap_add(d2i_trunc(khi), g.band + 1), 0, g.nk - 1);
    b.cj0 = iclamp(d2i_trunc(ceil(jlo)), 0, g.nj - 1);  b.cj1 = iclamp(d2i_trunc(floor(jhi)), 0, g.nj - 1);
    b.ck0 = iclamp(d2i_trunc(ceil(klo)), 0, g.nk - 1);  b.ck1 = iclamp(d2i_trunc(floor(khi)), 0, g.nk - 1);
    // clip to the planes this slab owns (clamps above use the GLOBAL nk)
    b.k0 = max(b.k0, g.k_lo);   b.k1 = min(b.k1, g.k_hi - 1);
    b.ck0 = max(b.ck0, g.k_lo); b.ck1 = min(b.ck1, g.k_hi - 1);
    // An extent can come out inverted along ANY axis: d2i_trunc yields INT_MIN for coordinates outside the int range
    // (|x - o| / dx >= 2^31, inf, NaN), so e.g. i0 = clamp(INT_MIN - band -> wraps) = ni-1 while i1 = 0.  The reference's
    // `for (i = i0; i <= i1; ++i)` loops (cpu_lib/makelevelset3.cpp:213-215) then run zero times; canonicalise such a box
    // to "k1 < k0", the one emptiness test band_voxels() and k_band make.
    if (b.i1 < b.i0 || b.j1 < b.j0) b.k1 = b.k0 - 1;
}

__device__ __forceinline__ uint32_t units_of(uint64_t n) { return (uint32_t)((n + UNIT - 1) / UNIT); }

// units[t] = band units + crossing units of triangle t; the integer extents are kept for k_band (the nine fp64
// divisions behind them cost more than the 48 bytes)
__global__ void k_count_units(const TriRec *__restrict__ rec, uint64_t ntri, Grid g, uint32_t *__restrict__ units,
                              TriExt *__restrict__ ext)
{
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntri) return;
    TriBoxes b;
    tri_boxes(rec[t], g, b);
    units[t] = units_of(b.band_voxels()) + units_of(b.cross_points());
    TriExt e;
    e.i0 = b.i0; e.i1 = b.i1; e.j0 = b.j0; e.j1 = b.j1; e.k0 = b.k0; e.k1 = b.k1;
    e.cj0 = b.cj0; e.cj1 = b.cj1; e.ck0 = b.ck0; e.ck1 = b.ck1; e.pad0 = 0; e.pad1 = 0;
    ext[t] = e;
}

// ---- exclusive prefix sum of uint32 -> uint64, three passes ------------------------------------
__device__ __forceinline__ uint64_t block_reduce_u64(uint64_t v, uint64_t *sh)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (l < (int)(blockDim.x >> 5)) ? sh[l] : 0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (l == 0) sh[0] = v;
    }
    __syncthreads();
    uint64_t r = sh[0];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(256) k_scan_reduce(const uint32_t *__restrict__ in, uint64_t n, uint64_t *__restrict__ block_sums)
{
    __shared__ uint64_t sh[32];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    uint64_t s = 0;
    for (int e = threadIdx.x; e < SCAN_TILE; e += 256) { uint64_t i = base + e; if (i < n) s += in[i]; }
    s = block_reduce_u64(s, sh);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = s;
}

// single block: in-place exclusive scan of block_sums[0..nb), total written to block_sums[nb]
__global__ void __launch_bounds__(1024) k_scan_block_sums(uint64_t *block_sums, uint32_t nb)
{
    __shared__ uint64_t sh[32];
    __shared__ uint64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint64_t v = (i < nb) ? block_sums[i] : 0;
        // inclusive warp scan
        uint64_t x = v;
        int l = threadIdx.x & 31, w = threadIdx.x >> 5;
        for (int o = 1; o < 32; o <<= 1) { uint64_t y = __shfl_up_sync(0xffffffffu, x, o); if (l >= o) x += y; }
        if (l == 31) sh[w] = x;
        __syncthreads();
        if (w == 0) {
            uint64_t s = sh[l];
            for (int o = 1; o < 32; o <<= 1) { uint64_t y = __shfl_up_sync(0xffffffffu, s, o); if (l >= o) s += y; }
            sh[l] = s;
        }
        __syncthreads();
        uint64_t warp_off = (w > 0) ? sh[w - 1] : 0;
        uint64_t carry = carry_s;
        if (i < nb) block_sums[i] = carry + warp_off + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_off + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[nb] = carry_s;
}

__global__ void __launch_bounds__(256) k_scan_final(const uint32_t *__restrict__ in, uint64_t n,
                                                    const uint64_t *__restrict__ block_sums, uint32_t nb,
                                                    uint64_t *__restrict__ prefix)
{
    __shared__ uint64_t sh[8];
    __shared__ uint64_t carry_s;
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    if (threadIdx.x == 0) carry_s = block_sums[blockIdx.x];
    __syncthreads();
    int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int e0 = 0; e0 < SCAN_TILE; e0 += 256) {
        uint64_t i = base + e0 + threadIdx.x;
        uint64_t v = (i < n) ? in[i] : 0;
        uint64_t x = v;
        for (int o = 1; o < 32; o <<= 1) { uint64_t y = __shfl_up_sync(0xffffffffu, x, o); if (l >= o) x += y; }
        if (l == 31) sh[w] = x;
        __syncthreads();
        uint64_t warp_off = 0;
        for (int u = 0; u < w; ++u) warp_off += sh[u];
        uint64_t carry = carry_s;
        if (i < n) prefix[i] = carry + warp_off + x - v;
        __syncthreads();
        if (threadIdx.x == 255) carry_s = carry + warp_off + x;
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) prefix[n] = block_sums[nb];
}

// ---- the band + crossing-count kernel ----------------------------------------------------------
// Persistent grid; each warp takes a CONTIGUOUS range of work units, so the triangle of a unit is found by
// one binary search per warp (prefix[t] <= unit < prefix[t+1]) and then by walking forward.
#ifndef SDFB_BAND_MINB
#define SDFB_BAND_MINB 4
#endif
__global__ void __launch_bounds__(256, SDFB_BAND_MINB) k_band(const TriRec *__restrict__ rec, const TriExt *__restrict__ ext, uint64_t ntri, Grid g,
                                              const uint64_t *__restrict__ prefix,
                                              uint64_t *__restrict__ cells, int32_t *__restrict__ counts,
                                              float init_phi)
{
    const int lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t wid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t total = prefix[ntri];
    const uint64_t per = (total + warps - 1) / warps;
    const uint64_t u0 = wid * per, u1 = min(total, u0 + per);
    if (u0 >= u1) return;
    uint64_t t;
    {
        uint64_t lo = 0, hi = ntri;          // invariant: prefix[lo] <= u0 < prefix[hi]
        while (hi - lo > 1) {
            uint64_t mid = (lo + hi) >> 1;
            if (__ldg(&prefix[mid]) <= u0) lo = mid; else hi = mid;
        }
        t = lo;
    }
    uint64_t t_end = __ldg(&prefix[t + 1]);  // first unit of the next triangle
    for (uint64_t unit = u0; unit < u1; ++unit) {
        while (t_end <= unit) { ++t; t_end = __ldg(&prefix[t + 1]); }      // skips triangles without units
        const uint32_t local = (uint32_t)(unit - __ldg(&prefix[t]));
        const TriRec tr = rec[t];
        const TriExt e = ext[t];
        const uint64_t nvox = (e.k1 < e.k0) ? 0 : (uint64_t)(e.i1 - e.i0 + 1) * (uint64_t)(e.j1 - e.j0 + 1) * (uint64_t)(e.k1 - e.k0 + 1);
        const uint32_t nband = units_of(nvox);
        if (local < nband) {
            // exact band: voxels [local*UNIT, ...) of the box, i fastest so a warp touches few lines
            const uint32_t wi = e.i1 - e.i0 + 1, wj = e.j1 - e.j0 + 1;
            const uint64_t v0 = (uint64_t)local * UNIT;
            const uint64_t v1 = min(nvox, v0 + UNIT);
            const bool small = nvox <= 0xffffffffull;   // 32-bit div/mod for all but absurd boxes
            for (uint64_t v = v0 + lane; v < v1; v += 32) {
                uint32_t i, j, k;
                if (small) {
                    uint32_t vv = (uint32_t)v, r = vv / wi;
                    i = e.i0 + (vv - r * wi); k = r / wj; j = e.j0 + (r - k * wj); k += e.k0;
                } else {
                    uint64_t r = v / wi;
                    i = e.i0 + (uint32_t)(v - r * wi); k = (uint32_t)(r / wj); j = e.j0 + (uint32_t)(r - (uint64_t)k * wj); k += e.k0;
                }
                F3 gx{lattice(i, g.dx, g.ox), lattice(j, g.dx, g.oy), lattice(k, g.dx, g.oz)};
                float d = ptd_rec(gx, tr);
                if (d < init_phi) {                       // also rejects NaN (degenerate triangles)
                    uint64_t cand = pack_cell(d, (uint32_t)t);
                    uint64_t *c = &cells[g.cidx(i, j, k)];
                    if (cand < *c) atomicMin(reinterpret_cast<unsigned long long *>(c), (unsigned long long)cand);
                }
            }
        } else {
            // x-ray crossings: lattice points [ (local-nband)*UNIT, ... ) of the yz range
            const uint64_t npts = (e.ck1 < e.ck0 || e.cj1 < e.cj0) ? 0 : (uint64_t)(e.cj1 - e.cj0 + 1) * (uint64_t)(e.ck1 - e.ck0 + 1);
            const uint32_t wj = e.cj1 - e.cj0 + 1;
            const uint64_t a0 = (uint64_t)(local - nband) * UNIT;
            const uint64_t a1 = min(npts, a0 + UNIT);
            const double fip = grid_coord(tr.p.x, g.ox, g.dx), fjp = grid_coord(tr.p.y, g.oy, g.dx), fkp = grid_coord(tr.p.z, g.oz, g.dx);
            const double fiq = grid_coord(tr.q.x, g.ox, g.dx), fjq = grid_coord(tr.q.y, g.oy, g.dx), fkq = grid_coord(tr.q.z, g.oz, g.dx);
            const double fir = grid_coord(tr.r.x, g.ox, g.dx), fjr = grid_coord(tr.r.y, g.oy, g.dx), fkr = grid_coord(tr.r.z, g.oz, g.dx);
            for (uint64_t a = a0 + lane; a < a1; a += 32) {
                int j = e.cj0 + (int)(a % wj);
                int k = e.ck0 + (int)(a / wj);
                double ba, bb, bc;
                if (point_in_triangle_2d((double)j, (double)k, fjp, fkp, fjq, fkq, fjr, fkr, ba, bb, bc)) {
                    double fi = __dadd_rn(__dadd_rn(__dmul_rn(ba, fip), __dmul_rn(bb, fiq)), __dmul_rn(bc, fir));
                    int ii = d2i_trunc(ceil(fi));
                    if (ii < 0) atomicAdd(&counts[g.vidx(0, j, k)], 1);
                    else if (ii < g.ni) atomicAdd(&counts[g.vidx(ii, j, k)], 1);
                }
            }
        }
    }
}

}  // namespace

int launch_init(uint64_t *cells, int64_t ncells, float init_phi, cudaStream_t st)
{
    uint32_t bits;
    memcpy(&bits, &init_phi, 4);
    uint64_t init_cell = ((uint64_t)bits << 32) | 0xffffffffu;
    k_init_cells<<<148 * 8, 256, 0, st>>>(cells, ncells, init_cell);
    return 1;
}

int launch_tri_prep(const uint32_t *tri, const float *xyz, uint64_t ntri, uint64_t nvert, TriRec *rec,
                    unsigned long long *bad, cudaStream_t st)
{
    if (ntri == 0) return 0;
    k_tri_prep<<<(unsigned)((ntri + 255) / 256), 256, 0, st>>>(tri, xyz, ntri, nvert, rec, bad);
    return 1;
}

int launch_band(const TriRec *rec, uint64_t ntri, const Grid &g, uint32_t *units, TriExt *ext, uint64_t *prefix,
                uint64_t *block_sums, uint64_t *cells, int32_t *counts, float init_phi, cudaStream_t st)
{
    if (ntri == 0) return 0;
    unsigned nb = (unsigned)((ntri + SCAN_TILE - 1) / SCAN_TILE);
    k_count_units<<<(unsigned)((ntri + 255) / 256), 256, 0, st>>>(rec, ntri, g, units, ext);
    k_scan_reduce<<<nb, 256, 0, st>>>(units, ntri, block_sums);
    k_scan_block_sums<<<1, 1024, 0, st>>>(block_sums, nb);
    k_scan_final<<<nb, 256, 0, st>>>(units, ntri, block_sums, nb, prefix);
    k_band<<<148 * SDFB_BAND_MINB * 2, 256, 0, st>>>(rec, ext, ntri, g, prefix, cells, counts, init_phi);
    return 5;
}

}  // namespace sdfb
