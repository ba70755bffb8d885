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

// Stamp memo table shared by the pipelined schedules.  last[c][m] = stamp (sweep index + 1) of the latest
// earlier sweep in which a voxel of class c examined neighbour offset m, 0 if none.  Class bits: 1 = last
// voxel of its row (ri = ni-1), 2 = last row (rj = nj-1), 4 = last plane (rk = nk-1): such voxels lie on a
// grid face and are only visited by sweeps whose direction along that axis equals the current one.  Offset m
// (cpu_lib/makelevelset3.cpp:143-149) steps along i for m in {0,2,4,6}, along j for {1,2,5,6}, along k for m >= 3.
inline void memo_last_table(int sweep_index, const SweepDir &cur, uint8_t (&last)[8][8])
{
    for (int c = 0; c < 8; ++c) for (int m = 0; m < 8; ++m) last[c][m] = 0;
    if (sweep_index + 1 > 31) return;                 // stamps saturate: no memo beyond 31 sweeps
    for (int c = 0; c < 8; ++c) for (int m = 0; m < 7; ++m) {
        const bool ci = (m == 0 || m == 2 || m == 4 || m == 6), cj = (m == 1 || m == 2 || m == 5 || m == 6), ck = (m >= 3);
        for (int e = sweep_index - 1; e >= 0; --e) {
            const SweepDir d = SweepDir::of(e);
            const bool same_i = d.di == cur.di, same_j = d.dj == cur.dj, same_k = d.dk == cur.dk;
            if ((!(ci || (c & 1)) || same_i) && (!(cj || (c & 2)) || same_j) && (!(ck || (c & 4)) || same_k)) { last[c][m] = (uint8_t)(e + 1); break; }
        }
    }
}

}  // namespace sdfb
