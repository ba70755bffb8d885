// sdfb_sweep_relax.cu -- sweep schedule "relax": the Gauss-Seidel sweep as a fixed-point iteration.
//
// A sweep of the reference (cpu_lib/makelevelset3.cpp:104-151) computes, for every voxel v in lexicographic
// order,   new[v] = G(old[v], new[n_0(v)], ..., new[n_6(v)])
// where n_m are the seven upstream neighbours (:143-149) and G folds check_neighbour (:90-102) over them in
// the reference's order with its strict "<".  The neighbours are upstream in a DAG, so this system of
// equations has exactly ONE solution -- the serial result -- and any iteration that (a) always evaluates G
// from the voxel's value at the START of the sweep (old[v], never from an intermediate value: an
// intermediate triangle may be closer than anything the serial order would ever show the voxel) and (b)
// re-evaluates a voxel after any of its seven inputs changed, ends in that solution.  That removes the
// wavefront: round 0 evaluates ALL voxels in parallel against whatever their neighbours currently hold,
// round r+1 re-evaluates only the downstream neighbours of the voxels that changed in round r.
//
// The cost is re-evaluation, so this schedule is for sweeps in which few voxels change: the second pass of
// the reference's two (0.005 .. 0.02 % of the voxels change per sweep at 512^3; the wavefront schedules still
// pay the full dependency depth ni+nj+nk for them).  Round 0 is then a streaming pass over the cells plus the
// distance evaluations the stamp memo could not exclude, all at full occupancy; the later rounds touch a few
// thousand voxels.  One cooperative launch per sweep: round 0 (dense), then rounds over a work list with a
// grid-wide barrier in between, until a round changes nothing.
//
// Bookkeeping: `oldbuf` (8 B per cell, touched only where a cell changes) keeps old[v] for voxels already
// changed in this sweep -- recognised by their stamp, which is this sweep's; work lists are de-duplicated
// with a bitmap (one bit per cell); a list that overflows falls back to scanning the bitmap.  The exact
// pruning rules (own / duplicate triangle, stamp memo) are those of sdfb_sweep_columns.cu.
#include <cstdio>
#include <cstdlib>
#include <cooperative_groups.h>
#include "sdfb_kernels.cuh"
#include "sdfb_sweep_common.cuh"

namespace cg = cooperative_groups;

namespace sdfb {

namespace {

constexpr int RX_WARPS = 8;                  // warps per CTA; in round 0 warp w handles row j0 + w
constexpr int RX_THREADS = RX_WARPS * 32;
constexpr int QCAP = 7 * 32;

struct RelaxParams {
    Grid g;
    SweepDir sd;
    int rk_first, rk_last;                   // relative k range updated by this launch (inclusive)
    uint32_t stamp;                          // sweep_index + 1 (< 31)
    uint32_t list_cap;                       // entries per work list
    uint64_t *cells;
    uint64_t *oldbuf;
    const TriRec *rec;
    uint32_t *list[2];                       // cell indices to re-evaluate, by round parity
    uint32_t *bitmap[2];                     // one bit per cell: "is in the list of that parity"
    unsigned int *count;                     // [3] list lengths, rotating by round % 3
    unsigned long long *debug;               // SDFB_RELAX_DEBUG: {ns round 0, ns total, round-1 list length, rounds}
    unsigned long long *changed;             // [0] cells whose triangle changed (net), [1] distance evaluations
    uint8_t last[8][8];
};

struct RelaxShared {
    uint32_t q_ent[RX_WARPS][QCAP];          // (owner lane << 27) | tri
    float q_d[RX_WARPS][QCAP];
    float px[RX_WARPS][32], py[RX_WARPS][32], pz[RX_WARPS][32];   // world position of each lane's voxel
    uint32_t thr[8][8];
    uint32_t tmin[8];                        // lowest threshold of each class: the cheap "nothing is fresh" test
};

__device__ __forceinline__ uint64_t ld_cg64(const uint64_t *p) { return __ldcg(reinterpret_cast<const unsigned long long *>(p)); }
__device__ __forceinline__ uint32_t ld_cg32(const uint64_t *cell) { return __ldcg(reinterpret_cast<const uint32_t *>(cell)); }

// Re-evaluates one voxel per lane (all 32 lanes must call; `valid` masks lanes without a voxel).
// Cells are read through L2: other SMs rewrite them during the launch.
__device__ __forceinline__ void relax_voxels(const RelaxParams &P, RelaxShared &sh, int warp, int lane, bool valid,
                                             int64_t c, int ri, int rj, int rk, int push_parity, unsigned int *push_count,
                                             int &net_changed, unsigned &evals)
{
    const Grid &g = P.g;
    uint32_t *const q_ent = sh.q_ent[warp];
    float *const q_d = sh.q_d[warp];
    const int64_t si = -(int64_t)P.sd.di, sj = -(int64_t)P.sd.dj * g.ni, sk = -(int64_t)P.sd.dk * g.plane();
    uint64_t cur64 = 0, base = 0;
    uint32_t nb[7];
    uint32_t live = 0;
    bool was_changed = false;
    if (valid) {
        const uint64_t *cp = P.cells + c;
        cur64 = ld_cg64(cp);
        nb[0] = ld_cg32(cp + si); nb[1] = ld_cg32(cp + sj); nb[2] = ld_cg32(cp + si + sj); nb[3] = ld_cg32(cp + sk);
        nb[4] = ld_cg32(cp + si + sk); nb[5] = ld_cg32(cp + sj + sk); nb[6] = ld_cg32(cp + si + sj + sk);
        const uint64_t old64 = ld_cg64(P.oldbuf + c);                 // speculative: only meaningful if the cell changed in this sweep
        was_changed = lo_stamp(cell_lo(cur64)) == P.stamp;
        base = was_changed ? old64 : cur64;
        const int cls = (ri == g.ni - 1 ? 1 : 0) | (rj == g.nj - 1 ? 2 : 0) | (rk == g.nk - 1 ? 4 : 0);
        const uint32_t own = cell_lo(base);
        const uint32_t mx = max(max(max(nb[0], nb[1]), max(nb[2], nb[3])), max(max(nb[4], nb[5]), nb[6]));
        if (mx >= sh.tmin[cls]) {
            #pragma unroll
            for (int m = 0; m < 7; ++m) {
                const uint32_t x = nb[m];
                const bool keep = ((x & TRI_MASK) != TRI_NONE) && (((x ^ own) & TRI_MASK) != 0) && (x >= sh.thr[cls][m]);
                live |= keep ? (1u << m) : 0u;
            }
        }
        if (live) {                           // drop repeats of ANY earlier neighbour's triangle
            #pragma unroll
            for (int m = 1; m < 7; ++m) {
                bool dup = false;
                #pragma unroll
                for (int u = 0; u < m; ++u) dup = dup || (((nb[u] ^ nb[m]) & TRI_MASK) == 0);
                if (dup) live &= ~(1u << m);
            }
        }
    }
    const int ncand = __popc(live);
    const uint32_t b0 = __ballot_sync(0xffffffffu, ncand & 1), b1 = __ballot_sync(0xffffffffu, ncand & 2),
                   b2 = __ballot_sync(0xffffffffu, ncand & 4);
    // a voxel changed earlier in this sweep must be re-derived even without candidates (it may have to revert)
    if ((b0 | b1 | b2) == 0 && !__any_sync(0xffffffffu, was_changed)) return;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int total = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
    const int off = __popc(b0 & lt_mask) + 2 * __popc(b1 & lt_mask) + 4 * __popc(b2 & lt_mask);
    if (live) {
        sh.px[warp][lane] = lattice(P.sd.abs_i(ri, g), g.dx, g.ox);
        sh.py[warp][lane] = lattice(P.sd.abs_j(rj, g), g.dx, g.oy);
        sh.pz[warp][lane] = lattice(P.sd.abs_k(rk, g), g.dx, g.oz);
        int q = off;
        #pragma unroll
        for (int m = 0; m < 7; ++m) if ((live >> m) & 1u) { q_ent[q] = ((uint32_t)lane << 27) | (nb[m] & TRI_MASK); ++q; }
    }
    __syncwarp();
    for (int q = lane; q < total; q += 32) {
        const uint32_t e = q_ent[q];
        const int ol = (int)(e >> 27);
        const F3 x0{sh.px[warp][ol], sh.py[warp][ol], sh.pz[warp][ol]};
        const TriRec *tr = &P.rec[e & TRI_MASK];
        const float4 p = __ldg(&tr->p), qq = __ldg(&tr->q), r = __ldg(&tr->r);
        q_d[q] = ptd_rec(x0, p, qq, r);
        ++evals;
    }
    __syncwarp();
    if (valid) {
        float phi = cell_phi(base);
        uint32_t best = TRI_NONE;
        for (int q = off; q < off + ncand; ++q) {                     // the reference's order and strict "<"
            const float d = q_d[q];
            if (d < phi) { phi = d; best = q_ent[q] & TRI_MASK; }
        }
        const uint64_t new64 = (best != TRI_NONE) ? pack_cell(phi, (P.stamp << 27) | best) : base;
        if (new64 != cur64) {
            if (!was_changed) P.oldbuf[c] = cur64;                    // == base: the value at the start of the sweep
            P.cells[c] = new64;
            net_changed += (best != TRI_NONE ? 1 : 0) - (was_changed ? 1 : 0);
            // schedule the (up to seven) downstream neighbours that this launch updates
            const bool pi = ri + 1 <= g.ni - 1, pj = rj + 1 <= g.nj - 1, pk = rk + 1 <= P.rk_last;
            uint32_t fresh = 0;                                       // bit m: neighbour m was not yet scheduled
            #pragma unroll
            for (int m = 1; m < 8; ++m) {                             // the atomics are independent: all in flight at once
                const bool a = m & 1, b = m & 2, cc = m & 4;
                if ((a && !pi) || (b && !pj) || (cc && !pk)) continue;
                const int64_t d = c - (a ? si : 0) - (b ? sj : 0) - (cc ? sk : 0);
                const uint32_t bit = 1u << (d & 31);
                const uint32_t prev = atomicOr(&P.bitmap[push_parity][d >> 5], bit);
                fresh |= (prev & bit) ? 0u : (1u << m);
            }
            if (fresh) {
                unsigned idx = atomicAdd(push_count, (unsigned)__popc(fresh));
                #pragma unroll
                for (int m = 1; m < 8; ++m) if ((fresh >> m) & 1u) {
                    const bool a = m & 1, b = m & 2, cc = m & 4;
                    const int64_t d = c - (a ? si : 0) - (b ? sj : 0) - (cc ? sk : 0);
                    if (idx < P.list_cap) P.list[push_parity][idx] = (uint32_t)d;
                    ++idx;
                }
            }
        }
    }
    __syncwarp();
}

// ---- kernel 1: streaming scan --------------------------------------------------------------------------------
// Marks (bitmap of round 0) every voxel that has at least one neighbour triangle to evaluate: one thread per
// voxel, lanes along i, 8 rows per CTA so that the rows a CTA shares are served by L1.  Nothing is written to
// the cells here, so the non-coherent path is fine.
constexpr int SCAN_ROWS = 8;
__global__ void __launch_bounds__(SCAN_ROWS * 32) k_relax_scan(RelaxParams P)
{
    __shared__ uint32_t thr[8][8];
    __shared__ uint32_t tmin[8];
    const Grid &g = P.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 64) {
        const uint32_t l = P.last[tid >> 3][tid & 7];
        thr[tid >> 3][tid & 7] = ((tid & 7) == 7) ? 0xffffffffu : (l ? (l + 1u) << 27 : 0u);
    }
    __syncthreads();
    if (tid < 8) {
        uint32_t t = 0xffffffffu;
        for (int m = 0; m < 7; ++m) t = min(t, thr[tid][m]);
        tmin[tid] = t;
    }
    __syncthreads();
    const int rj = 1 + blockIdx.x * SCAN_ROWS + warp, rk = P.rk_first + blockIdx.y;
    if (rj > g.nj - 1) return;                                        // warp-uniform
    const int64_t row = g.cidx(0, P.sd.abs_j(rj, g), P.sd.abs_k(rk, g));
    const int64_t si = -(int64_t)P.sd.di, sj = -(int64_t)P.sd.dj * g.ni, sk = -(int64_t)P.sd.dk * g.plane();
    const int cls_row = (rj == g.nj - 1 ? 2 : 0) | (rk == g.nk - 1 ? 4 : 0);
    bool any_work = false;
    #pragma unroll 4
    for (int i0 = 0; i0 < g.ni; i0 += 32) {
        const int i = i0 + lane;
        const int ri = P.sd.di > 0 ? i : g.ni - 1 - i;
        bool any_live = false;
        if (i < g.ni && ri >= 1) {
            const uint32_t *cp = reinterpret_cast<const uint32_t *>(P.cells + row + i);     // low words: {stamp | tri}
            uint32_t nb[7];
            const uint32_t own = __ldg(cp);
            nb[0] = __ldg(cp + 2 * si); nb[1] = __ldg(cp + 2 * sj); nb[2] = __ldg(cp + 2 * (si + sj)); nb[3] = __ldg(cp + 2 * sk);
            nb[4] = __ldg(cp + 2 * (si + sk)); nb[5] = __ldg(cp + 2 * (sj + sk)); nb[6] = __ldg(cp + 2 * (si + sj + sk));
            const int cls = cls_row | (ri == g.ni - 1 ? 1 : 0);
            const uint32_t mx = max(max(max(nb[0], nb[1]), max(nb[2], nb[3])), max(max(nb[4], nb[5]), nb[6]));
            if (mx >= tmin[cls]) {
                // (repeats of an earlier neighbour's triangle are not removed here: marking too much is harmless)
                #pragma unroll
                for (int m = 0; m < 7; ++m) {
                    const uint32_t x = nb[m];
                    any_live = any_live || (((x & TRI_MASK) != TRI_NONE) && (((x ^ own) & TRI_MASK) != 0) && (x >= thr[cls][m]));
                }
            }
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, any_live);
        if (bal) {
            const int64_t c0 = row + i0;
            const int shf = (int)(c0 & 31);
            if (lane == 0) atomicOr(&P.bitmap[0][c0 >> 5], bal << shf);
            if (lane == 1 && shf && (bal >> (32 - shf))) atomicOr(&P.bitmap[0][(c0 >> 5) + 1], bal >> (32 - shf));
            any_work = true;
        }
    }
    if (any_work && lane == 0) *reinterpret_cast<volatile unsigned int *>(&P.count[0]) = 1u;          // "round 0 has work"
}

// ---- kernel 2: the rounds ---------------------------------------------------------------------------------------
constexpr unsigned SOLO_MAX = 512;           // lists this short are finished by one CTA (a CTA barrier per round
                                             // instead of a grid barrier)

#ifndef SDFB_RELAX_MINB
#define SDFB_RELAX_MINB 3
#endif
__global__ void __launch_bounds__(RX_THREADS, SDFB_RELAX_MINB) k_relax_rounds(RelaxParams P)
{
    __shared__ RelaxShared sh;
    cg::grid_group grid = cg::this_grid();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 64) {
        const uint32_t l = P.last[tid >> 3][tid & 7];
        sh.thr[tid >> 3][tid & 7] = ((tid & 7) == 7) ? 0xffffffffu : (l ? (l + 1u) << 27 : 0u);
    }
    __syncthreads();
    if (tid < 8) {
        uint32_t t = 0xffffffffu;
        for (int m = 0; m < 7; ++m) t = min(t, sh.thr[tid][m]);
        sh.tmin[tid] = t;
    }
    __syncthreads();
    int net_changed = 0;
    unsigned evals = 0;
    unsigned long long t_start = 0;
    if (P.debug && blockIdx.x == 0 && tid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_start));
    const int64_t gwarp = (int64_t)blockIdx.x * RX_WARPS + warp, nwarps = (int64_t)gridDim.x * RX_WARPS;

    // Round r reads the list of parity r&1 (length count[r%3]) and fills the other one (count[(r+1)%3]);
    // count[(r+2)%3] was consumed in round r-1 and is refilled in round r+1: reset it now.  Once a list is
    // short, CTA 0 finishes alone (a CTA barrier per round instead of a grid barrier); a list that grows again
    // is still handled correctly, just by that CTA.
    const Grid &g = P.g;
    const int64_t plane = g.plane();
    const int64_t nwords = (g.cell_count() + 31) >> 5;
    int r = 0;
    bool solo = false;
    for (;; ++r) {
        const unsigned n = *reinterpret_cast<volatile unsigned int *>(&P.count[r % 3]);
        if (n == 0) break;
        if (!solo && r > 0 && n <= SOLO_MAX) {                        // uniform over the grid
            solo = true;
            if (blockIdx.x != 0) break;
        }
        if (blockIdx.x == 0 && tid == 0) P.count[(r + 2) % 3] = 0;
        unsigned int *const push_count = &P.count[(r + 1) % 3];
        const int par = r & 1;
        // the work list: the bitmap in round 0 (filled by the scan kernel) and when a list overflowed, else the list
        const bool use_bitmap = (r == 0) || n > P.list_cap;
        const int64_t w = solo ? warp : gwarp, nw = solo ? RX_WARPS : nwarps;
        const int64_t limit = use_bitmap ? nwords : (int64_t)n;
        for (int64_t pos = w * 32; pos < limit; pos += nw * 32) {
            uint32_t mine = 0, nz = 1;
            if (use_bitmap) {                                         // 32 words per warp: each set word = 32 consecutive cells
                if (pos + lane < nwords) {
                    mine = __ldcg(&P.bitmap[par][pos + lane]);
                    if (mine) P.bitmap[par][pos + lane] = 0;          // pushes of this round go to the other parity
                }
                nz = __ballot_sync(0xffffffffu, mine != 0);
            }
            while (nz) {
                const int src = __ffs(nz) - 1;
                nz &= nz - 1;
                bool valid;
                int64_t c = 0;
                if (use_bitmap) {
                    const uint32_t bits = __shfl_sync(0xffffffffu, mine, src);
                    valid = (bits >> lane) & 1u;
                    c = (pos + src) * 32 + lane;
                } else {
                    valid = pos + lane < (int64_t)n;
                    if (valid) {
                        c = (int64_t)__ldcg(&P.list[par][pos + lane]);
                        atomicAnd(&P.bitmap[par][c >> 5], ~(1u << (c & 31)));
                    }
                }
                int ri = 0, rj = 0, rk = 0;
                if (valid) {
                    const int64_t p = c / plane, rem = c - p * plane;
                    const int j = (int)(rem / g.ni), i = (int)(rem - (int64_t)j * g.ni), k = (int)p - 1 + g.k_lo;
                    ri = P.sd.di > 0 ? i : g.ni - 1 - i;
                    rj = P.sd.dj > 0 ? j : g.nj - 1 - j;
                    rk = P.sd.rel_k(k, g);
                }
                relax_voxels(P, sh, warp, lane, valid, c, ri, rj, rk, par ^ 1, push_count, net_changed, evals);
            }
        }
        if (solo) __syncthreads();                                    // orders the CTA's global writes and reads
        else grid.sync();
        if (r == 0 && P.debug && blockIdx.x == 0 && tid == 0) {
            unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            P.debug[0] = t - t_start; P.debug[2] = P.count[1];
        }
    }
    if (P.debug && blockIdx.x == 0 && tid == 0) {
        unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        P.debug[1] = t - t_start; P.debug[3] = (unsigned long long)r;
    }

    // ---- teardown ---------------------------------------------------------------------------------------------
    for (int o = 16; o > 0; o >>= 1) { net_changed += __shfl_down_sync(0xffffffffu, net_changed, o); evals += __shfl_down_sync(0xffffffffu, evals, o); }
    if (lane == 0 && net_changed) atomicAdd(P.changed, (unsigned long long)(long long)net_changed);
    if (lane == 0 && evals) atomicAdd(P.changed + 1, (unsigned long long)evals);
}

}  // namespace

// scratch the schedule needs for a slab: oldbuf (8 B per cell), two lists, two bitmaps, three counters
bool sweep_relax_supported(const Grid &g) { return g.cell_count() < ((int64_t)1 << 32); }
uint32_t sweep_relax_list_cap(const Grid &g)
{
    const int64_t cap = g.cell_count() < ((int64_t)16 << 20) ? g.cell_count() : ((int64_t)16 << 20);
    return (uint32_t)cap;
}
size_t sweep_relax_scratch_bytes(const Grid &g)
{
    const size_t cells = (size_t)g.cell_count(), words = (cells + 31) / 32 + 1;
    return cells * 8 + 2 * (size_t)sweep_relax_list_cap(g) * 4 + 2 * words * 4 + 64;
}

// `scratch` must be zero-initialised once after allocation (bitmaps and counters return to zero after every sweep).
int launch_sweep_relax(uint64_t *cells, const TriRec *rec, const Grid &g, int sweep_index,
                       unsigned long long *changed, void *scratch, cudaStream_t st)
{
    RelaxParams P{};
    P.g = g;
    P.sd = SweepDir::of(sweep_index);
    int rk_lo, rk_hi;
    if (!P.sd.owned_rk_range(g, rk_lo, rk_hi)) return 0;
    if (g.ni < 2 || g.nj < 2) return 0;
    P.rk_first = rk_lo; P.rk_last = rk_hi;
    P.stamp = (uint32_t)(sweep_index + 1);
    P.cells = cells; P.rec = rec; P.changed = changed;
    P.list_cap = sweep_relax_list_cap(g);
    const size_t ncells = (size_t)g.cell_count(), words = (ncells + 31) / 32 + 1;
    char *s = static_cast<char *>(scratch);
    P.count = reinterpret_cast<unsigned int *>(s); s += 64;
    P.oldbuf = reinterpret_cast<uint64_t *>(s); s += ncells * 8;
    P.list[0] = reinterpret_cast<uint32_t *>(s); s += (size_t)P.list_cap * 4;
    P.list[1] = reinterpret_cast<uint32_t *>(s); s += (size_t)P.list_cap * 4;
    P.bitmap[0] = reinterpret_cast<uint32_t *>(s); s += words * 4;
    P.bitmap[1] = reinterpret_cast<uint32_t *>(s);
    memo_last_table(sweep_index, P.sd, P.last);
    cudaMemsetAsync(P.count, 0, 64, st);
    int dev = 0, sms = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_relax_rounds, RX_THREADS, 0);
    if (occ < 1) occ = 1;
    static unsigned long long *dbg = nullptr;
    if (getenv("SDFB_RELAX_DEBUG")) {
        if (!dbg) cudaMalloc(&dbg, 4 * sizeof(unsigned long long));
        cudaMemsetAsync(dbg, 0, 4 * sizeof(unsigned long long), st);
        P.debug = dbg;
    }
    const dim3 sgrid((g.nj - 1 + SCAN_ROWS - 1) / SCAN_ROWS, rk_hi - rk_lo + 1);
    k_relax_scan<<<sgrid, SCAN_ROWS * 32, 0, st>>>(P);
    void *args[] = {&P};
    cudaLaunchCooperativeKernel((const void *)k_relax_rounds, dim3(sms * occ), dim3(RX_THREADS), args, 0, st);
    if (P.debug) {
        unsigned long long h[4];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[relax] sweep %2d: round 0 %.3f ms, total %.3f ms, first list %llu, rounds %llu, grid %d x %d\n", sweep_index,
                h[0] * 1e-6, h[1] * 1e-6, h[2], h[3], sms * occ, RX_THREADS);
    }
    return 2;
}

}  // namespace sdfb
