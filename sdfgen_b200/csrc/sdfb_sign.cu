// sdfb_sign.cu -- phase C of make_level_set3 on sm_100a: inside/outside sign from the running parity
// of x-ray crossing counts along i (reference: cpu_lib/makelevelset3.cpp:295-303), fused with the
// unpacking of the 8-byte cells into the float output, plus the output layout helpers.
//
// Rows (j,k) are independent and contiguous in i, so one warp owns a row and walks it 32 voxels at a
// time: the parity prefix inside a chunk is a ballot + popc, the carry between chunks is one bit.
#include "sdfb_kernels.cuh"

namespace sdfb {

namespace {

__global__ void __launch_bounds__(256) k_sign_rows(const uint64_t *__restrict__ cells, const int32_t *__restrict__ counts,
                                                   Grid g, int apply_sign, float *__restrict__ phi_out)
{
    const int lane = threadIdx.x & 31;
    const int64_t rows = (int64_t)g.nj * g.nkl();
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps) {
        const int64_t vbase = row * g.ni;                    // slab-local dense index of (0,j,k)
        const int64_t cbase = vbase + g.plane();             // cells have one halo plane in front
        unsigned carry = 0;
        for (int i0 = 0; i0 < g.ni; i0 += 32) {
            int i = i0 + lane;
            bool in = i < g.ni;
            unsigned odd = 0;
            float phi = 0.f;
            if (in) {
                phi = cell_phi(cells[cbase + i]);
                if (apply_sign) odd = (unsigned)counts[vbase + i] & 1u;
            }
            unsigned bal = __ballot_sync(0xffffffffu, odd);
            unsigned par = (__popc(bal & (0xffffffffu >> (31 - lane))) + carry) & 1u;   // inclusive prefix
            if (in) phi_out[vbase + i] = par ? -phi : phi;
            carry = (carry + __popc(bal)) & 1u;
        }
    }
}

__global__ void __launch_bounds__(256) k_unpack_tri(const uint64_t *__restrict__ cells, int64_t n, int64_t halo,
                                                    int32_t *__restrict__ tri_out)
{
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride)
        tri_out[v] = lo_tri_signed(cell_lo(cells[halo + v]));
}

// i-fastest [k][j][i]  ->  k-fastest [i][j][k] for one slab; 32x32 (i,k) tiles per j through shared memory
__global__ void __launch_bounds__(256) k_relayout(const uint32_t *__restrict__ src, int ni, int nj, int nk,
                                                  uint32_t *__restrict__ dst)
{
    __shared__ uint32_t tile[32][33];
    const int j = blockIdx.z;
    const int i0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        int i = i0 + tx, k = k0 + r;
        if (i < ni && k < nk) tile[r][tx] = src[(int64_t)i + (int64_t)ni * ((int64_t)j + (int64_t)nj * k)];
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        int i = i0 + r, k = k0 + tx;
        if (i < ni && k < nk) dst[((int64_t)i * nj + j) * (int64_t)nk + k] = tile[tx][r];
    }
}

// number of values < 0 (the .sdf writer's "inside" count, common/sdf_io.cpp:53: -0.0f is not inside)
__global__ void __launch_bounds__(256) k_count_negative(const float *__restrict__ v, int64_t n, unsigned long long *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned mine = 0;
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += stride) mine += v[x] < 0.0f ? 1u : 0u;
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(out, (unsigned long long)mine);
}

// Halo planes received from a neighbouring slab: raise the stamp to the maximum so that the sweep's
// "unchanged since I last looked" memo never skips them (the copy this slab looked at may have been stale).
__global__ void __launch_bounds__(256) k_halo_refresh(uint64_t *__restrict__ cells, int64_t plane, int64_t far_off)
{
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < 2 * plane; v += stride) {
        uint64_t *c = cells + (v < plane ? v : far_off + (v - plane));
        uint64_t x = *c;
        if ((cell_lo(x) & TRI_MASK) != TRI_NONE) *c = x | ((uint64_t)31u << 27);
    }
}

// Verification pass (checks for grids the reference cannot run): every cell must hold EXACTLY the
// reference distance to the triangle it names (phi == point_triangle_distance(gx, tri[closest_tri]), bit for bit), or the
// initial distance where no triangle was ever assigned.  Also folds the slab into two order-independent 64-bit
// checksums keyed by the GLOBAL voxel index, so that the checksums of the slabs of a sharded run add up (mod 2^64) to
// the checksum of one plan on the whole grid: out[2] over whole cell words (distance, stamp, triangle), out[3] over
// (distance, triangle) only.  out[0] = inconsistent cells, out[1] = cells without a triangle.
__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256) k_verify_cells(const uint64_t *__restrict__ cells, const TriRec *__restrict__ rec, Grid g,
                                                      uint32_t init_bits, unsigned long long *__restrict__ out)
{
    const int64_t n = g.slab_voxels(), stride = (int64_t)gridDim.x * blockDim.x, plane = g.plane();
    unsigned long long bad = 0, none = 0, sum_cell = 0, sum_val = 0;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
        const uint64_t c = cells[plane + v];
        const uint32_t lo = cell_lo(c), t = lo_tri(lo), pb = (uint32_t)(c >> 32);
        const int64_t kk = v / plane, r = v - kk * plane;
        const int j = (int)(r / g.ni), i = (int)(r - (int64_t)j * g.ni), k = (int)kk + g.k_lo;
        if (t == TRI_NONE) { ++none; bad += pb != init_bits; }
        else {
            const F3 gx{lattice(i, g.dx, g.ox), lattice(j, g.dx, g.oy), lattice(k, g.dx, g.oz)};
            bad += __float_as_uint(ptd_rec(gx, rec[t])) != pb;
        }
        const uint64_t gidx = (uint64_t)v + (uint64_t)g.k_lo * (uint64_t)plane;
        sum_cell += mix64(mix64(gidx) ^ c);
        sum_val += mix64(mix64(gidx) ^ (((uint64_t)pb << 32) | t));
    }
    for (int o = 16; o > 0; o >>= 1) {
        bad += __shfl_down_sync(0xffffffffu, bad, o); none += __shfl_down_sync(0xffffffffu, none, o);
        sum_cell += __shfl_down_sync(0xffffffffu, sum_cell, o); sum_val += __shfl_down_sync(0xffffffffu, sum_val, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (bad) atomicAdd(out, bad);
        if (none) atomicAdd(out + 1, none);
        atomicAdd(out + 2, sum_cell);
        atomicAdd(out + 3, sum_val);
    }
}

}  // namespace

int launch_verify_cells(const uint64_t *cells, const TriRec *rec, const Grid &g, float init_phi, unsigned long long *out, cudaStream_t st)
{
    k_verify_cells<<<148 * 8, 256, 0, st>>>(cells, rec, g, __builtin_bit_cast(uint32_t, init_phi), out);
    return 1;
}

int launch_halo_refresh(uint64_t *cells, const Grid &g, cudaStream_t st)
{
    k_halo_refresh<<<148 * 2, 256, 0, st>>>(cells, g.plane(), g.plane() * (int64_t)(g.nkl() + 1));
    return 1;
}

int launch_count_negative(const float *v, int64_t n, unsigned long long *out, cudaStream_t st)
{
    k_count_negative<<<148 * 8, 256, 0, st>>>(v, n, out);
    return 1;
}

int launch_sign(const uint64_t *cells, const int32_t *counts, const Grid &g, bool apply_sign,
                bool kfastest, float *phi_out, cudaStream_t st)
{
    (void)kfastest;   // layout change is a separate launch_relayout_i32 so this kernel stays streaming
    k_sign_rows<<<148 * 8, 256, 0, st>>>(cells, counts, g, apply_sign ? 1 : 0, phi_out);
    return 1;
}

int launch_unpack_tri(const uint64_t *cells, const Grid &g, bool kfastest, int32_t *tri_out, cudaStream_t st)
{
    (void)kfastest;
    k_unpack_tri<<<148 * 8, 256, 0, st>>>(cells, g.slab_voxels(), g.plane(), tri_out);
    return 1;
}

int launch_relayout_i32(const int32_t *src, const Grid &g, int32_t *dst_kfastest, cudaStream_t st)
{
    dim3 grid((g.ni + 31) / 32, (g.nkl() + 31) / 32, g.nj);
    k_relayout<<<grid, 256, 0, st>>>(reinterpret_cast<const uint32_t *>(src), g.ni, g.nj, g.nkl(),
                                     reinterpret_cast<uint32_t *>(dst_kfastest));
    return 1;
}

}  // namespace sdfb
