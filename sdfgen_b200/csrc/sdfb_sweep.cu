// sdfb_sweep.cu -- phase B of make_level_set3 on sm_100a: the 2 x 8 closest-triangle sweeps
// (reference: cpu_lib/makelevelset3.cpp:90-151 and the driver loop :243-292, one thread).
//
// A sweep is a Gauss-Seidel pass: in sweep-relative coordinates (r = index counted from the corner
// the sweep starts at) voxel (ri,rj,rk), ri,rj,rk >= 1, looks at the closest triangle of its seven
// upstream neighbours (ri-1|ri, rj-1|rj, rk-1|rk) in the fixed order of :143-149 and adopts one if
// it is strictly closer.  All seven lie on earlier anti-diagonals ri+rj+rk, so any schedule that
// respects that partial order reproduces the serial result bit for bit.
//
// Two schedules are provided:
//   levels   (SDFB_SWEEP_LEVELS) one launch per anti-diagonal; trivially exact, launch-bound; kept
//            as the on-device cross-check for grids the CPU oracle cannot finish.
//   columns  (default) see sdfb_sweep_columns.cu.
#include "sdfb_kernels.cuh"
#include "sdfb_sweep_common.cuh"

namespace sdfb {

namespace {

// One anti-diagonal level w = ri+rj+rk of one sweep.  Thread <-> (rj, rk); ri follows.
__global__ void __launch_bounds__(256) k_sweep_level(uint64_t *__restrict__ cells, const TriRec *__restrict__ rec,
                                                     Grid g, SweepDir sd, int w, int rk_first, int nrk,
                                                     uint32_t stamp, unsigned long long *__restrict__ changed)
{
    int rj = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    int rk = rk_first + blockIdx.y;
    int ri = w - rj - rk;
    bool active = (rj < g.nj) && (blockIdx.y < (unsigned)nrk) && (ri >= 1) && (ri < g.ni);
    bool did = false;
    if (active) {
        int i = sd.abs_i(ri, g), j = sd.abs_j(rj, g), k = sd.abs_k(rk, g);
        int64_t c = g.cidx(i, j, k);
        int64_t si = -(int64_t)sd.di, sj = -(int64_t)sd.dj * g.ni, sk = -(int64_t)sd.dk * g.plane();
        uint64_t self = cells[c];
        uint32_t nb[7];
        nb[0] = cell_lo(cells[c + si]);
        nb[1] = cell_lo(cells[c + sj]);
        nb[2] = cell_lo(cells[c + si + sj]);
        nb[3] = cell_lo(cells[c + sk]);
        nb[4] = cell_lo(cells[c + si + sk]);
        nb[5] = cell_lo(cells[c + sj + sk]);
        nb[6] = cell_lo(cells[c + si + sj + sk]);
        F3 gx{lattice(i, g.dx, g.ox), lattice(j, g.dx, g.oy), lattice(k, g.dx, g.oz)};
        float phi = cell_phi(self);
        uint32_t cur = lo_tri(cell_lo(self));
        #pragma unroll 1
        for (int m = 0; m < 7; ++m) {
            uint32_t t = lo_tri(nb[m]);
            if (t == TRI_NONE || t == cur) continue;      // same triangle: d == phi, never "<"
            float d = ptd_rec(gx, rec[t]);
            if (d < phi) { phi = d; cur = t; did = true; }
        }
        if (did) cells[c] = pack_cell(phi, (stamp << 27) | cur);
    }
    unsigned m = __ballot_sync(0xffffffffu, did);
    if (m && (threadIdx.x & 31) == 0) atomicAdd(changed, (unsigned long long)__popc(m));
}

}  // namespace

int launch_sweep_levels(uint64_t *cells, const TriRec *rec, const Grid &g, int sweep_index,
                        unsigned long long *changed, cudaStream_t st)
{
    SweepDir sd = SweepDir::of(sweep_index);
    int rk_lo, rk_hi;
    if (!sd.owned_rk_range(g, rk_lo, rk_hi)) return 0;
    if (g.ni < 2 || g.nj < 2) return 0;
    int n = 0;
    int nrk = rk_hi - rk_lo + 1;
    dim3 block(256), grid((g.nj - 1 + 255) / 256, nrk);
    for (int w = 2 + rk_lo; w <= (g.ni - 1) + (g.nj - 1) + rk_hi; ++w) {
        // planes that can hold voxels of this level: 1 <= w - rj - rk <= ni-1 with 1 <= rj <= nj-1
        int a = max(rk_lo, w - (g.ni - 1) - (g.nj - 1)), b = min(rk_hi, w - 2);
        if (a > b) continue;
        grid.y = b - a + 1;
        k_sweep_level<<<grid, block, 0, st>>>(cells, rec, g, sd, w, a, b - a + 1, (uint32_t)(sweep_index + 1), changed);
        ++n;
    }
    return n;
}

}  // namespace sdfb
