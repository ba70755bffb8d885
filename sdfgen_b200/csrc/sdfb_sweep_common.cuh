// sdfb_sweep_common.cuh -- sweep directions and sweep-relative coordinates.
#pragma once
#include "sdfb_kernels.cuh"

namespace sdfb {

// Direction s%8 of the reference's table, cpu_lib/makelevelset3.cpp:245-248.
struct SweepDir {
    int di, dj, dk;
    __host__ __device__ static SweepDir of(int s)
    {
        // {+,+,+},{-,-,-},{+,+,-},{-,-,+},{+,-,+},{-,+,-},{+,-,-},{-,+,+}
        const int q = (s % 8) >> 1;           // pair index: 0:(++ +) 1:(+ + -) 2:(+ - +) 3:(+ - -)
        const int flip = (s & 1) ? -1 : 1;    // odd entries are the mirrored direction
        const int bj = (q & 2) ? -1 : 1, bk = (q & 1) ? -1 : 1;
        return SweepDir{flip, flip * bj, flip * bk};
    }
    // sweep-relative index r (distance from the face the sweep starts at) -> absolute index.
    // The reference visits r = 1 .. n-1 (i0=1 or n-2, :109-116); r = 0 is only ever read.
    __host__ __device__ int abs_i(int r, const Grid &g) const { return di > 0 ? r : g.ni - 1 - r; }
    __host__ __device__ int abs_j(int r, const Grid &g) const { return dj > 0 ? r : g.nj - 1 - r; }
    __host__ __device__ int abs_k(int r, const Grid &g) const { return dk > 0 ? r : g.nk - 1 - r; }
    __host__ __device__ int rel_k(int k, const Grid &g) const { return dk > 0 ? k : g.nk - 1 - k; }
    // relative k range [lo,hi] of the planes this slab owns AND the sweep updates (rk >= 1); false if empty
    __host__ __device__ bool owned_rk_range(const Grid &g, int &lo, int &hi) const
    {
        int a = rel_k(g.k_lo, g), b = rel_k(g.k_hi - 1, g);
        lo = a < b ? a : b; hi = a < b ? b : a;
        if (lo < 1) lo = 1;
        return lo <= hi;
    }
};

}  // namespace sdfb
