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
// grid-wide barrier in between, until a round changes nothing; once a list is short, one CTA finishes alone with
// the lists in shared memory.  For the last sweeps of the pass, which have almost no candidates, a lean scan
// kernel marks the voxels that have one and round 0 starts from that bitmap instead of the dense pass.
// Evaluations are batched per warp (filter as the voxels come, evaluate a full queue with all lanes, replay).
// oracle/relax_emu.c runs the same rules on the CPU in random order against the serial oracle.
//
// Bookkeeping: `oldbuf` (8 B per cell, touched only where a cell changes) keeps old[v] for voxels already
// changed in this sweep -- recognised by their stamp, which is this sweep's; work lists are de-duplicated
// with a bitmap (one bit per cell); a list that overflows falls back to scanning the bitmap.  The exact
// pruning rules (own / duplicate triangle, stamp memo) are those of sdfb_sweep_columns.cu.
#include <cstdio>
#include <cstdlib>
#include "sdfb_kernels.cuh"
#include "sdfb_sweep_common.cuh"

namespace sdfb {

namespace {

constexpr int RX_WARPS = 8;                  // warps per CTA; in round 0 warp w handles row j0 + w
constexpr int RX_THREADS = RX_WARPS * 32;

struct RelaxParams {
    Grid g;
    SweepDir sd;
    int rk_first, rk_last;                   // relative k range updated by this launch (inclusive)
    uint32_t stamp;                          // sweep_index + 1 (< 31)
    uint32_t list_cap;                       // entries per work list
    int scan_mode;                           // round 0 = streaming scan kernel + bitmap instead of the dense pass
    uint64_t *cells;
    uint64_t *oldbuf;
    const TriRec *rec;
    uint32_t *list[2];                       // cell indices to re-evaluate, by round parity
    uint32_t *bitmap[2];                     // one bit per cell: "is in the list of that parity"
    unsigned int *count;                     // [0..2] list lengths, rotating by round % 3; [4] grid barrier arrivals
    uint32_t heavy_limit;                    // more work-list entries than this in all: give the sweep back to the column schedule
    unsigned long long *debug;               // SDFB_RELAX_DEBUG: {ns round 0, ns total, round-1 list length, rounds}
    unsigned long long *changed;             // [0] cells whose triangle changed (net), [1] distance evaluations
    uint8_t last[8][8];
};

// Per-warp batch: voxels are filtered as they come (one per lane and call), their candidate triangles are
// appended to the warp's queue, and the queue is evaluated when it is nearly full -- with all 32 lanes busy,
// whatever the number of candidates per voxel -- before each pending voxel replays its own results.
#ifndef SDFB_RELAX_QCAP
#define SDFB_RELAX_QCAP 416
#endif
#ifndef SDFB_RELAX_PCAP
#define SDFB_RELAX_PCAP 96
#endif
#ifndef SDFB_RELAX_SOLOCAP
#define SDFB_RELAX_SOLOCAP 2048
#endif
constexpr int QCAP_B = SDFB_RELAX_QCAP;                  // queue entries per warp; a call adds at most 7 * 32
constexpr int PCAP_B = SDFB_RELAX_PCAP;                   // pending voxels per warp; a call adds at most 32
constexpr int SOLO_CAP = SDFB_RELAX_SOLOCAP;               // entries of the in-CTA work lists of the tail rounds
struct Pending {
    uint32_t c;                              // cell index
    float px, py, pz;                        // world position
    uint32_t info;                           // queue offset (16 bits) | candidates << 16 | may-push flags i,j,k << 20 | was_changed << 23
    uint32_t cur_lo, cur_hi;                 // the cell as the filter read it (nobody else writes it in this round)
    uint32_t base_lo, base_hi;               // the cell at the start of the sweep
};
struct RelaxShared {
    uint32_t q_ent[RX_WARPS][QCAP_B];        // triangle
    float q_d[RX_WARPS][QCAP_B];             // owner (index into pend) until evaluated, then the distance
    Pending pend[RX_WARPS][PCAP_B];
    uint32_t thr[8][8];
    uint32_t tmin[8];                        // lowest threshold of each class: the cheap "nothing is fresh" test
    uint32_t slist[2][SOLO_CAP];             // tail rounds (one CTA): the work lists live here
    unsigned int scount[3];
};

__device__ __forceinline__ uint64_t ld_cg64(const uint64_t *p) { return __ldcg(reinterpret_cast<const unsigned long long *>(p)); }
__device__ __forceinline__ uint32_t ld_cg32(const uint64_t *cell) { return __ldcg(reinterpret_cast<const uint32_t *>(cell)); }

// Filter one voxel per lane (all 32 lanes must call; `valid` masks lanes without a voxel) and append it to the
// warp's batch.  own = the voxel's {stamp|tri} word at the START of the sweep, nb = its neighbours' current words,
// was_changed = the cell was already rewritten in this sweep.  nq / np = entries in the candidate queue / pending
// list (warp-uniform).
__device__ __forceinline__ void relax_filter(const RelaxParams &P, RelaxShared &sh, int warp, int lane, bool valid,
                                             int64_t c, int ri, int rj, int rk, uint64_t cur64, uint64_t base64,
                                             const uint32_t (&nb)[7], bool was_changed, int &nq, int &np)
{
    const Grid &g = P.g;
    const uint32_t own = cell_lo(base64);
    uint32_t live = 0;
    bool pend = false;
    if (valid) {
        const int cls = (ri == g.ni - 1 ? 1 : 0) | (rj == g.nj - 1 ? 2 : 0) | (rk == g.nk - 1 ? 4 : 0);
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
            uint32_t t[7];
            #pragma unroll
            for (int m = 0; m < 7; ++m) t[m] = nb[m] & TRI_MASK;
            #pragma unroll
            for (int m = 1; m < 7; ++m) {
                bool dup = false;
                #pragma unroll
                for (int u = 0; u < m; ++u) dup = dup || (t[u] == t[m]);
                if (dup) live &= ~(1u << m);
            }
        }
        // a voxel changed earlier in this sweep must be re-derived even without candidates (it may have to revert)
        pend = live || was_changed;
    }
    const uint32_t bp = __ballot_sync(0xffffffffu, pend);
    if (!bp) return;
    const int ncand = __popc(live);
    const uint32_t b0 = __ballot_sync(0xffffffffu, ncand & 1), b1 = __ballot_sync(0xffffffffu, ncand & 2),
                   b2 = __ballot_sync(0xffffffffu, ncand & 4);
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int off = nq + __popc(b0 & lt_mask) + 2 * __popc(b1 & lt_mask) + 4 * __popc(b2 & lt_mask);
    const int pidx = np + __popc(bp & lt_mask);
    if (pend) {
        Pending &pe = sh.pend[warp][pidx];
        pe.c = (uint32_t)c;
        pe.px = lattice(P.sd.abs_i(ri, g), g.dx, g.ox);
        pe.py = lattice(P.sd.abs_j(rj, g), g.dx, g.oy);
        pe.pz = lattice(P.sd.abs_k(rk, g), g.dx, g.oz);
        pe.info = (uint32_t)off | ((uint32_t)ncand << 16) | (ri + 1 <= g.ni - 1 ? 1u << 20 : 0u) | (rj + 1 <= g.nj - 1 ? 1u << 21 : 0u) |
                  (rk + 1 <= P.rk_last ? 1u << 22 : 0u) | (was_changed ? 1u << 23 : 0u);
        pe.cur_lo = cell_lo(cur64); pe.cur_hi = (uint32_t)(cur64 >> 32);
        pe.base_lo = cell_lo(base64); pe.base_hi = (uint32_t)(base64 >> 32);
        int q = off;
        #pragma unroll
        for (int m = 0; m < 7; ++m) if ((live >> m) & 1u) {
            sh.q_ent[warp][q] = nb[m] & TRI_MASK;
            sh.q_d[warp][q] = __int_as_float(pidx);
            ++q;
        }
    }
    nq += __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
    np += __popc(bp);
}

// The eight {stamp|tri} words a voxel's filter needs, through L1 (round 0: a value read too early is repaired by
// the re-evaluation its writer schedules) ...
struct Words { uint64_t own; uint32_t nb[7]; };
__device__ __forceinline__ void load_words_l1(const uint64_t *cp, int64_t si, int64_t sj, int64_t sk, Words &w)
{
    const uint32_t *p = reinterpret_cast<const uint32_t *>(cp);
    w.own = *cp;
    w.nb[0] = p[2 * si]; w.nb[1] = p[2 * sj]; w.nb[2] = p[2 * (si + sj)]; w.nb[3] = p[2 * sk];
    w.nb[4] = p[2 * (si + sk)]; w.nb[5] = p[2 * (sj + sk)]; w.nb[6] = p[2 * (si + sj + sk)];
}
// ... or through L2 (later rounds: other SMs rewrite cells during the launch).
__device__ __forceinline__ void load_words_l2(const uint64_t *cp, int64_t si, int64_t sj, int64_t sk, Words &w)
{
    w.own = ld_cg64(cp);
    w.nb[0] = ld_cg32(cp + si); w.nb[1] = ld_cg32(cp + sj); w.nb[2] = ld_cg32(cp + si + sj); w.nb[3] = ld_cg32(cp + sk);
    w.nb[4] = ld_cg32(cp + si + sk); w.nb[5] = ld_cg32(cp + sj + sk); w.nb[6] = ld_cg32(cp + si + sj + sk);
}

// Evaluate the warp's queue, then let every pending voxel replay its results in the reference's order.
// solo = tail rounds run by one CTA: the next list and its length live in shared memory (entries beyond
// SOLO_CAP spill to the global list).
__device__ __forceinline__ void relax_flush(const RelaxParams &P, RelaxShared &sh, int warp, int lane, int &nq, int &np,
                                            int push_parity, unsigned int *push_count, bool solo, int &net_changed, unsigned &evals)
{
    if (np == 0) return;
    const Grid &g = P.g;
    uint32_t *const q_ent = sh.q_ent[warp];
    float *const q_d = sh.q_d[warp];
    const Pending *const pend = sh.pend[warp];
    const int64_t si = -(int64_t)P.sd.di, sj = -(int64_t)P.sd.dj * g.ni, sk = -(int64_t)P.sd.dk * g.plane();
    __syncwarp();
    for (int q = lane; q < nq; q += 32) {
        const Pending &pe = pend[__float_as_int(q_d[q])];
        const F3 x0{pe.px, pe.py, pe.pz};
        const TriRec *tr = &P.rec[q_ent[q]];
        const float4 p = __ldg(&tr->p), qq = __ldg(&tr->q), r = __ldg(&tr->r);
        q_d[q] = ptd_rec(x0, p, qq, r);
        ++evals;
    }
    __syncwarp();
    for (int pi = lane; pi < np; pi += 32) {
        const Pending &pe = pend[pi];
        const int64_t c = (int64_t)pe.c;
        const uint32_t info = pe.info;
        const bool was_changed = (info >> 23) & 1u;
        const uint64_t cur64 = ((uint64_t)pe.cur_hi << 32) | pe.cur_lo, base = ((uint64_t)pe.base_hi << 32) | pe.base_lo;
        float phi = cell_phi(base);
        uint32_t best = TRI_NONE;
        const int q0 = (int)(info & 0xffffu), q1 = q0 + (int)((info >> 16) & 7u);
        for (int q = q0; q < q1; ++q) {                               // the reference's order and strict "<"
            const float d = q_d[q];
            if (d < phi) { phi = d; best = q_ent[q]; }
        }
        const uint64_t new64 = (best != TRI_NONE) ? pack_cell(phi, (P.stamp << 27) | best) : base;
        if (new64 != cur64) {
            if (!was_changed) P.oldbuf[c] = cur64;                    // == base: the value at the start of the sweep
            P.cells[c] = new64;
            net_changed += (best != TRI_NONE ? 1 : 0) - (was_changed ? 1 : 0);
            // schedule the (up to seven) downstream neighbours that this launch updates
            const bool pi_ok = (info >> 20) & 1u, pj_ok = (info >> 21) & 1u, pk_ok = (info >> 22) & 1u;
            uint32_t fresh = 0;                                       // bit m: neighbour m was not yet scheduled
            #pragma unroll
            for (int m = 1; m < 8; ++m) {                             // the atomics are independent: all in flight at once
                const bool a = m & 1, b = m & 2, cc = m & 4;
                if ((a && !pi_ok) || (b && !pj_ok) || (cc && !pk_ok)) continue;
                const int64_t d = c - (a ? si : 0) - (b ? sj : 0) - (cc ? sk : 0);
                const uint32_t bit = 1u << (d & 31);
                const uint32_t prev = atomicOr(&P.bitmap[push_parity][d >> 5], bit);
                fresh |= (prev & bit) ? 0u : (1u << m);
            }
            if (fresh) {
                unsigned idx = atomicAdd(push_count, (unsigned)__popc(fresh));     // shared memory in solo mode
                #pragma unroll
                for (int m = 1; m < 8; ++m) if ((fresh >> m) & 1u) {
                    const bool a = m & 1, b = m & 2, cc = m & 4;
                    const int64_t d = c - (a ? si : 0) - (b ? sj : 0) - (cc ? sk : 0);
                    if (solo && idx < (unsigned)SOLO_CAP) sh.slist[push_parity][idx] = (uint32_t)d;
                    else if (idx - (solo ? (unsigned)SOLO_CAP : 0u) < P.list_cap) P.list[push_parity][idx - (solo ? (unsigned)SOLO_CAP : 0u)] = (uint32_t)d;
                    ++idx;
                }
            }
        }
    }
    __syncwarp();
    nq = 0; np = 0;
}

// ---- optional kernel 0: streaming scan ------------------------------------------------------------------------
// For sweeps in which almost no voxel has a candidate, round 0 is split: this lean kernel (30 registers, full
// occupancy) streams over the cells and only MARKS the voxels that have at least one neighbour triangle to evaluate
// (bitmap of parity 0); the rounds kernel then starts from that bitmap.  One thread per voxel, lanes along i, 8 rows
// per CTA so that the rows a CTA shares are served by L1.  Nothing is written to the cells here.
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
    // (loading only the four rows' words at ri and taking the ri-1 ones from the lane below by shuffle measured slower)
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

// Grid-wide barrier for the co-resident (cooperatively launched) CTAs: one arrival counter that only grows;
// `target` is the value it reaches when every CTA has arrived at this barrier.  Several times cheaper than
// cooperative_groups' grid.sync() here, and the rounds are all latency.
__device__ __forceinline__ void grid_barrier(unsigned int *ctr, unsigned int &target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(ctr, 1u);
        while (*reinterpret_cast<volatile unsigned int *>(ctr) < target) { }
        __threadfence();
    }
    __syncthreads();
}

// ---- the kernel: round 0 over all voxels, then the rounds ---------------------------------------------------------------------------------------
constexpr int MAX_GRID_ROUNDS = 256;         // more grid-wide rounds than this cost as much as a column sweep
constexpr unsigned SOLO_MAX = 512;           // lists this short are finished by one CTA (a CTA barrier per round
                                             // instead of a grid barrier)

#ifndef SDFB_RELAX_MINB
#define SDFB_RELAX_MINB 3
#endif
__global__ void __launch_bounds__(RX_THREADS, SDFB_RELAX_MINB) k_relax_rounds(RelaxParams P)
{
    extern __shared__ __align__(16) unsigned char relax_smem[];
    RelaxShared &sh = *reinterpret_cast<RelaxShared *>(relax_smem);
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

    const Grid &g = P.g;
    const int64_t si = -(int64_t)P.sd.di, sj = -(int64_t)P.sd.dj * g.ni, sk = -(int64_t)P.sd.dk * g.plane();
    int nq = 0, np = 0;

    // ---- round 0: every voxel the sweep updates.  CTA = 8 consecutive rows of one plane (warp = row, lanes
    // along the row); the words of the next 32 voxels are loaded while the current ones are filtered. -------
    unsigned int bar_target = 0;
    if (!P.scan_mode) {
        const int nrows = g.nj - 1, nplanes = P.rk_last - P.rk_first + 1;
        const int jblocks = (nrows + RX_WARPS - 1) / RX_WARPS;
        const int64_t nitems = (int64_t)jblocks * nplanes;
        // each warp asks L2 for the row of its NEXT item (one bulk prefetch per 2 KB) while it works on the current
        // one: the few warps an SM holds cannot keep enough loads in flight to hide DRAM latency themselves
        auto prefetch_row = [&](int64_t item) {
            if (item >= nitems) return;
            const int prk = P.rk_first + (int)(item / jblocks), prj = 1 + (int)(item % jblocks) * RX_WARPS + warp;
            if (prj > g.nj - 1) return;
            const uintptr_t beg = reinterpret_cast<uintptr_t>(P.cells + g.cidx(0, P.sd.abs_j(prj, g), P.sd.abs_k(prk, g))) & ~(uintptr_t)15;
            const uint32_t bytes = ((uint32_t)g.ni * 8u + 16u) & ~15u;
            for (uint32_t o = (uint32_t)lane * 2048u; o < bytes; o += 32u * 2048u) {
                const uint32_t len = min(2048u, bytes - o);
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(beg + o), "r"(len) : "memory");
            }
        };
        prefetch_row(blockIdx.x);
        for (int64_t it = blockIdx.x; it < nitems; it += gridDim.x) {
            prefetch_row(it + gridDim.x);
            // a sweep that changes a large part of the grid is the column schedule's (see the fallback below):
            // the counter only grows, so every CTA agrees after the barrier whoever notices first
            if (*reinterpret_cast<volatile unsigned int *>(&P.count[1]) > P.heavy_limit) { nq = np = 0; break; }
            const int rk = P.rk_first + (int)(it / jblocks);
            const int rj = 1 + (int)(it % jblocks) * RX_WARPS + warp;
            if (rj > g.nj - 1) continue;                              // warp-uniform
            const uint64_t *row = P.cells + g.cidx(0, P.sd.abs_j(rj, g), P.sd.abs_k(rk, g));
            const int i_skip = P.sd.di > 0 ? 0 : g.ni - 1;            // the ri = 0 voxel is only read
            Words wa, wb;
            bool va = lane < g.ni && lane != i_skip, vb = false;
            if (va) load_words_l1(row + lane, si, sj, sk, wa);
            for (int i0 = 0; i0 < g.ni; i0 += 64) {
                {
                    const int i = i0 + 32 + lane;
                    vb = i < g.ni && i != i_skip;
                    if (vb) load_words_l1(row + i, si, sj, sk, wb);
                }
                {
                    const int i = i0 + lane;
                    relax_filter(P, sh, warp, lane, va, (row - P.cells) + i, P.sd.di > 0 ? i : g.ni - 1 - i, rj, rk, wa.own, wa.own, wa.nb, false, nq, np);
                    if (nq > QCAP_B - 7 * 32 || np > PCAP_B - 32) relax_flush(P, sh, warp, lane, nq, np, 1, &P.count[1], false, net_changed, evals);
                }
                if (i0 + 32 >= g.ni) break;
                {
                    const int i = i0 + 64 + lane;
                    va = i < g.ni && i != i_skip;
                    if (va) load_words_l1(row + i, si, sj, sk, wa);
                }
                {
                    const int i = i0 + 32 + lane;
                    relax_filter(P, sh, warp, lane, vb, (row - P.cells) + i, P.sd.di > 0 ? i : g.ni - 1 - i, rj, rk, wb.own, wb.own, wb.nb, false, nq, np);
                    if (nq > QCAP_B - 7 * 32 || np > PCAP_B - 32) relax_flush(P, sh, warp, lane, nq, np, 1, &P.count[1], false, net_changed, evals);
                }
            }
        }
        relax_flush(P, sh, warp, lane, nq, np, 1, &P.count[1], false, net_changed, evals);
        grid_barrier(&P.count[4], bar_target);
        if (P.debug && blockIdx.x == 0 && tid == 0) {
            unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            P.debug[0] = t - t_start; P.debug[2] = P.count[1];
        }
    }

    // ---- rounds 1, 2, ...: round r reads the list of parity r&1 (length count[r%3]) and fills the other one
    // (count[(r+1)%3]); count[(r+2)%3] was consumed in round r-1 and is refilled in round r+1: reset it now.
    // Once a list is short, CTA 0 finishes alone (a CTA barrier per round instead of a grid barrier); a list
    // that grows again is still handled correctly, just by that CTA. ------------------------------------------
    const uint32_t plane32 = (uint32_t)g.plane();
    const int64_t nwords = (g.cell_count() + 31) >> 5;
    int r = P.scan_mode ? 0 : 1;             // with the scan kernel, round 0 runs here from the bitmap it filled
    bool solo = false;
    unsigned long long work = 0;             // list entries so far (the same number in every CTA)
    for (;; ++r) {
        const int par = r & 1;
        unsigned n = solo ? *reinterpret_cast<volatile unsigned int *>(&sh.scount[r % 3])
                          : *reinterpret_cast<volatile unsigned int *>(&P.count[r % 3]);
        if (n == 0) break;
        work += n;
        if (!solo && r >= 1 && (work > P.heavy_limit || r > MAX_GRID_ROUNDS)) {                         // uniform over the grid
            // ---- fallback: this sweep changes too much for relaxation to pay (every changed voxel re-opens its
            // downstream neighbours, and a front that crosses an empty region is re-evaluated again and again:
            // such sweeps take seconds).  Put the cells back as they were before
            // the sweep (changed cells carry this sweep's stamp and their old value is in oldbuf), empty both
            // bitmaps and raise the flag that lets the conditional column launch queued behind this one run. -----
            // Owned planes only: in the exact multi-GPU mode a halo cell may carry this sweep's stamp too (the upstream
            // slab changed it in this very sweep) and has no oldbuf entry here.
            const int64_t c_end = g.plane() * (g.nkl() + 1), gt = (int64_t)blockIdx.x * RX_THREADS + tid, nt = (int64_t)gridDim.x * RX_THREADS;
            for (int64_t c = g.plane() + gt; c < c_end; c += nt) {
                const uint64_t x = ld_cg64(P.cells + c);
                if (lo_stamp(cell_lo(x)) == P.stamp) P.cells[c] = ld_cg64(P.oldbuf + c);
            }
            for (int64_t wi = gt; wi < nwords; wi += nt) { P.bitmap[0][wi] = 0; P.bitmap[1][wi] = 0; }
            if (blockIdx.x == 0 && tid == 0) P.count[5] = 1u;
            net_changed = 0;
            break;
        }
        if (P.debug && blockIdx.x == 0 && tid == 0 && r < 250) {
            unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            P.debug[4 + 2 * r] = n; P.debug[5 + 2 * r] = t - t_start;
        }
        if (!solo && r > 0 && n <= SOLO_MAX) {                        // uniform over the grid
            solo = true;
            if (blockIdx.x != 0) break;
            for (unsigned t = tid; t < n; t += RX_THREADS) sh.slist[par][t] = __ldcg(&P.list[par][t]);     // import the list
            if (tid == 0) sh.scount[(r + 1) % 3] = 0;
            __syncthreads();
        }
        if (solo) { if (tid == 0) sh.scount[(r + 2) % 3] = 0; }
        else if (blockIdx.x == 0 && tid == 0) P.count[(r + 2) % 3] = 0;
        unsigned int *const push_count = solo ? &sh.scount[(r + 1) % 3] : &P.count[(r + 1) % 3];
        // the bitmap is the work list when a list overflowed
        const bool use_bitmap = r == 0 || n > P.list_cap + (solo ? (unsigned)SOLO_CAP : 0u);
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
                        const int64_t idx = pos + lane;
                        if (solo) c = idx < SOLO_CAP ? (int64_t)sh.slist[par][idx] : (int64_t)__ldcg(&P.list[par][idx - SOLO_CAP]);
                        else c = (int64_t)__ldcg(&P.list[par][idx]);
                        atomicAnd(&P.bitmap[par][c >> 5], ~(1u << (c & 31)));
                    }
                }
                int ri = 0, rj = 0, rk = 0;
                Words wd;
                uint64_t base64 = 0;
                bool was_changed = false;
                if (valid) {                                          // cell indices fit 32 bits (sweep_relax_supported)
                    const uint32_t p = (uint32_t)c / plane32, rem = (uint32_t)c - p * plane32;
                    const int j = (int)(rem / (uint32_t)g.ni), i = (int)(rem - (uint32_t)j * (uint32_t)g.ni), k = (int)p - 1 + g.k_lo;
                    ri = P.sd.di > 0 ? i : g.ni - 1 - i;
                    rj = P.sd.dj > 0 ? j : g.nj - 1 - j;
                    rk = P.sd.rel_k(k, g);
                    load_words_l2(P.cells + c, si, sj, sk, wd);
                    const uint64_t old64 = ld_cg64(P.oldbuf + c);     // speculative: only meaningful if the cell changed in this sweep
                    was_changed = lo_stamp(cell_lo(wd.own)) == P.stamp;
                    base64 = was_changed ? old64 : wd.own;
                }
                relax_filter(P, sh, warp, lane, valid, c, ri, rj, rk, wd.own, base64, wd.nb, was_changed, nq, np);
                if (nq > QCAP_B - 7 * 32 || np > PCAP_B - 32) relax_flush(P, sh, warp, lane, nq, np, par ^ 1, push_count, solo, net_changed, evals);
            }
        }
        relax_flush(P, sh, warp, lane, nq, np, par ^ 1, push_count, solo, net_changed, evals);
        if (solo) __syncthreads();                                    // orders the CTA's writes (global and shared) and reads
        else grid_barrier(&P.count[4], bar_target);
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

// word that is non-zero after a launch that gave its sweep back (launch_sweep_columns' run_if)
const unsigned int *sweep_relax_fallback_flag(const void *scratch) { return static_cast<const unsigned int *>(scratch) + 5; }

// `scratch` must be zero-initialised once after allocation (bitmaps and counters return to zero after every sweep).
int launch_sweep_relax(uint64_t *cells, const TriRec *rec, const Grid &g, int sweep_index,
                       unsigned long long *changed, void *scratch, cudaStream_t st, const Tuning &tun, int max_ctas)
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
    if (tun.relax_list_cap > 0) P.list_cap = min(P.list_cap, (uint32_t)tun.relax_list_cap);   // tests: force the bitmap fallback
    const size_t ncells = (size_t)g.cell_count(), words = (ncells + 31) / 32 + 1;
    char *s = static_cast<char *>(scratch);
    P.count = reinterpret_cast<unsigned int *>(s); s += 64;
    P.oldbuf = reinterpret_cast<uint64_t *>(s); s += ncells * 8;
    P.list[0] = reinterpret_cast<uint32_t *>(s); s += (size_t)P.list_cap * 4;
    P.list[1] = reinterpret_cast<uint32_t *>(s); s += (size_t)P.list_cap * 4;
    P.bitmap[0] = reinterpret_cast<uint32_t *>(s); s += words * 4;
    P.bitmap[1] = reinterpret_cast<uint32_t *>(s);
    // light sweeps put well under 1 % of the cells on their work lists (C2: 0.004-0.1 % on the first, less after),
    // heavy ones most of them, round after round
    P.heavy_limit = (uint32_t)(ncells / 64);
    if (tun.relax_heavy_limit >= 0) P.heavy_limit = (uint32_t)tun.relax_heavy_limit;   // tests: force the fallback
    memo_last_table(sweep_index, P.sd, P.last);
    cudaMemsetAsync(P.count, 0, 64, st);
    int dev = 0, sms = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = sizeof(RelaxShared);
    static int occ_cached[64] = {0};          // per device: the attribute and the occupancy query are per-launch overhead
    if (dev < 0 || dev >= 64 || !occ_cached[dev]) {
        cudaFuncSetAttribute(k_relax_rounds, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_relax_rounds, RX_THREADS, smem);
        if (occ < 1) occ = 1;
        if (dev >= 0 && dev < 64) occ_cached[dev] = occ;
    } else occ = occ_cached[dev];
    static unsigned long long *dbg = nullptr;
    if (tun.relax_debug) {
        if (!dbg) cudaMalloc(&dbg, 512 * sizeof(unsigned long long));
        cudaMemsetAsync(dbg, 0, 512 * sizeof(unsigned long long), st);
        P.debug = dbg;
    }
    // sweeps late in the second pass have almost no candidates: stream over the cells with the lean scan kernel
    P.scan_mode = sweep_index >= tun.relax_scan_from ? 1 : 0;
    int launches = 1;
    if (P.scan_mode) {
        const dim3 sgrid((g.nj - 1 + SCAN_ROWS - 1) / SCAN_ROWS, rk_hi - rk_lo + 1);
        k_relax_scan<<<sgrid, SCAN_ROWS * 32, 0, st>>>(P);
        ++launches;
    }
    void *args[] = {&P};
    int grid = sms * occ;                                            // all CTAs of a cooperative launch are co-resident
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;            // batch mode: several plans share the device
    cudaLaunchCooperativeKernel((const void *)k_relax_rounds, dim3(grid), dim3(RX_THREADS), args, smem, st);
    if (P.debug) {
        unsigned long long h[512];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
        if (tun.relax_debug > 1) {
            fprintf(stderr, "[relax] sweep %2d rounds (n @ us):", sweep_index);
            for (int r = 1; r < 250 && h[4 + 2 * r]; ++r) fprintf(stderr, " %llu@%.0f", h[4 + 2 * r], h[5 + 2 * r] * 1e-3);
            fprintf(stderr, "\n");
        }
        fprintf(stderr, "[relax] sweep %2d: round 0 %.3f ms, total %.3f ms, first list %llu, rounds %llu, grid %d x %d\n", sweep_index,
                h[0] * 1e-6, h[1] * 1e-6, h[2], h[3], grid, RX_THREADS);
    }
    return launches;
}

}  // namespace sdfb
