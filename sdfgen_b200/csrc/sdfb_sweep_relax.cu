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
// grid-wide barrier in between, until a round changes nothing; lists of at most 4096 entries are run by a team of
// 16 CTAs, lists of at most 256 by one CTA with the lists (and the set of cells already pushed) in shared memory.
// Evaluations are batched per warp (filter as the voxels come, evaluate a full queue with all lanes, replay).
// oracle/relax_emu.c runs the same rules on the CPU in random order against the serial oracle.
//
// Round 0 has two cheaper forms.  (1) The lookahead window (below; the default from sweep 8 on for whole grids): ONE scan
// of the cells serves up to eight consecutive sweeps, each of which then starts from a short list.  (2) Without a
// window, the last sweeps of a pass, which have almost no candidates, start from a bitmap that a lean scan kernel fills.
//
// Bookkeeping: `oldbuf` (8 B per cell, touched only where a cell changes) keeps old[v] for voxels already
// changed in this sweep -- recognised by their stamp, which is this sweep's; work lists are de-duplicated
// with a bitmap (one bit per cell); a list that overflows falls back to scanning the bitmap.  The exact
// pruning rules (own / duplicate triangle, stamp memo) are those of sdfb_sweep_columns.cu.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "sdfb_kernels.cuh"
#include "sdfb_sweep_common.cuh"

namespace sdfb {

namespace {

constexpr int RX_WARPS = 8;                  // warps per CTA; in round 0 warp w handles row j0 + w
constexpr int RX_THREADS = RX_WARPS * 32;

// ---- lookahead over the sweeps of a pass (second pass and later) ------------------------------------------------------
// In those sweeps 0.005 .. 0.02 % of the voxels change, yet every sweep's round 0 streams over all cells and evaluates
// whatever the stamp memo cannot exclude (0.6 .. 0.001 evaluations per voxel) only to find that nearly all of them lose.
// k_look_scan does that work for up to eight consecutive sweeps in ONE pass over the cells as they are before the first
// of them: for every voxel and every sweep s of the window it evaluates the candidates sweep s would see IF nothing
// around the voxel changed in between, and records the voxel in w_list[s % 8] when one of them beats the voxel's
// distance.  Sweep s then starts from  W_s  u  D_s,  D_s = the cells changed by the window's earlier sweeps (c_list) and
// their seven downstream neighbours in sweep s's direction: a voxel outside both has exactly the candidates the scan
// evaluated (same neighbour words, same thresholds -- the memo tables of the window's own sweeps already exclude what
// an earlier sweep of the window looked at) against the same distance, and they all lost; everything else is the
// unchanged relaxation (rounds, re-evaluation of downstream neighbours, reverts).  Marking too much is harmless.
// Any event that makes the lists incomplete (a list overflows, a sweep is handed back to the column schedule) raises
// dense_off, after which the remaining sweeps of the window do their own dense round 0 as before.
struct LookState {
    unsigned int dense_off;                  // lookahead data unusable for the rest of the window
    unsigned int c_count;                    // entries in c_list (may run past cap_c: then dense_off is set)
    unsigned int w_count[8];                 // entries in w_list[q]
    unsigned int pad[6];
};

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
    unsigned int *count;                     // [0..2] list lengths, rotating by round % 3; [4] grid barrier arrivals; [5] "handed back to the
                                             // columns"; [6] team barrier arrivals; [7] "round 0 starts from the lookahead lists" (k_look_mark)
    uint32_t heavy_limit;                    // more work-list entries than this in all: give the sweep back to the column schedule
    unsigned long long *debug;               // SDFB_RELAX_DEBUG: {ns round 0, ns total, round-1 list length, rounds}
    unsigned long long *changed;             // [0] cells whose triangle changed (net), [1] distance evaluations
    uint8_t last[8][8];
    // lookahead window (k_look_scan below): when set and not switched off on the device, round 0 starts from the bitmap
    // k_look_mark filled, and every cell this sweep changes is appended to c_list
    LookState *look;
    uint32_t *c_list;                        // cells changed by the sweeps of the window so far (duplicates allowed)
    uint32_t cap_c;
    uint32_t *w_list;                        // [8][cap_w]: per direction, voxels with a winning candidate in the state before the window
    uint32_t cap_w;
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
#define SDFB_RELAX_SOLOCAP 1024
#endif
#ifndef SDFB_RELAX_PF
#define SDFB_RELAX_PF 1
#endif
constexpr int QCAP_B = SDFB_RELAX_QCAP;                  // queue entries per warp; a call adds at most 7 * 32
constexpr int PCAP_B = SDFB_RELAX_PCAP;                   // pending voxels per warp; a call adds at most 32
constexpr int SOLO_CAP = SDFB_RELAX_SOLOCAP;               // entries of the in-CTA work lists of the tail rounds
// Single-CTA rounds of at most HASH_MAX entries de-duplicate their pushes in shared memory instead of with atomicOr on the
// global bitmap: waiting for those seven atomics was 54 % of such a round (profiles/r2_lookahead.txt).  The table is emptied
// at the start of the round and receives at most 7 * HASH_MAX of its HSET slots, so a probe always ends.
constexpr int HSET = 2048;
constexpr unsigned HASH_MAX = 256;
constexpr uint32_t HSET_EMPTY = 0xffffffffu;               // no cell index (sweep_relax_supported: fewer than 2^32 cells)
struct Pending {
    uint32_t c;                              // cell index
    float px, py, pz;                        // world position
    uint32_t info;                           // queue offset (16 bits) | candidates << 16 | may-push flags i,j,k << 20 | was_changed << 23
    uint32_t cur_lo, cur_hi;                 // the cell as the filter read it (nobody else writes it in this round)
    uint32_t base_lo, base_hi;               // the cell at the start of the sweep
};
// -DSDFB_RELAX_PHASES (measurement builds): cycles of the single-CTA rounds by phase, warp 0 of CTA 0, summed into
// debug[300 ..]: 4 = list + cell loads, 0 = filter, 1 = evaluation, 2 = replay + pushes, 3 = CTA barrier + next length
#ifdef SDFB_RELAX_PHASES
#define RELAX_PHASE(k) do { if (solo && P.debug && warp == 0 && lane == 0) { const unsigned long long t_ = clock64(); \
        atomicAdd(&P.debug[300 + (k)], t_ - sh.tprev); sh.tprev = t_; } } while (0)
#else
#define RELAX_PHASE(k) do { } while (0)
#endif
struct RelaxShared {
    uint32_t q_ent[RX_WARPS][QCAP_B];        // triangle
    float q_d[RX_WARPS][QCAP_B];             // owner (index into pend) until evaluated, then the distance
    Pending pend[RX_WARPS][PCAP_B];
    uint32_t thr[8][8];
    uint32_t tmin[8];                        // lowest threshold of each class: the cheap "nothing is fresh" test
    uint32_t slist[2][SOLO_CAP];             // tail rounds (one CTA): the work lists live here
    uint32_t hset[HSET];                     // ... and the set of cells already pushed in this round (open addressing)
    unsigned int scount[3];
    unsigned long long tprev;                // SDFB_RELAX_PHASES
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
                                            int push_parity, unsigned int *push_count, bool solo, int &net_changed, unsigned &evals,
                                            bool use_hash = false)
{
    if (np == 0) return;
    const Grid &g = P.g;
    uint32_t *const q_ent = sh.q_ent[warp];
    float *const q_d = sh.q_d[warp];
    const Pending *const pend = sh.pend[warp];
    const int64_t si = -(int64_t)P.sd.di, sj = -(int64_t)P.sd.dj * g.ni, sk = -(int64_t)P.sd.dk * g.plane();
    __syncwarp();
    RELAX_PHASE(0);
    for (int q = lane; q < nq; q += 32) {
        const Pending &pe = pend[__float_as_int(q_d[q])];
        const F3 x0{pe.px, pe.py, pe.pz};
        const TriRec *tr = &P.rec[q_ent[q]];
        const float4 p = __ldg(&tr->p), qq = __ldg(&tr->q), r = __ldg(&tr->r);
        q_d[q] = ptd_rec(x0, p, qq, r);
        ++evals;
    }
    __syncwarp();
    RELAX_PHASE(1);
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
            if (!was_changed) {
                P.oldbuf[c] = cur64;                                  // == base: the value at the start of the sweep
                if (P.look) {                                         // lookahead window: later sweeps start from what changed
                    const unsigned idx = atomicAdd(&P.look->c_count, 1u);
                    if (idx < P.cap_c) P.c_list[idx] = (uint32_t)c;
                    else P.look->dense_off = 1u;
                }
            }
            P.cells[c] = new64;
            net_changed += (best != TRI_NONE ? 1 : 0) - (was_changed ? 1 : 0);
            // The rounds are chains of dependent memory round trips.  The next round re-evaluates this voxel's downstream
            // neighbours: their cells and the cells around them (the 3 x 3 rows around c) and their oldbuf entries are what it
            // will wait for -- ask L2 for them now (SDFB_RELAX_PF=0 builds without).
#if SDFB_RELAX_PF
            {
                const int64_t ncell = g.cell_count(), rj = (int64_t)g.ni, rk = g.plane();
                #pragma unroll
                for (int ok = -1; ok <= 1; ++ok) {
                    #pragma unroll
                    for (int oj = -1; oj <= 1; ++oj) {
                        const int64_t a = c + oj * rj + ok * rk;
                        if (a >= 0 && a < ncell) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.cells + a));
                    }
                }
                #pragma unroll
                for (int m = 1; m < 4; ++m) {                         // oldbuf rows of the downstream neighbours (the row of c was just written)
                    const int64_t a = c - ((m & 1) ? sj : 0) - ((m & 2) ? sk : 0);
                    if (a >= 0 && a < ncell) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.oldbuf + a));
                }
            }
#endif
            // schedule the (up to seven) downstream neighbours that this launch updates
            const bool pi_ok = (info >> 20) & 1u, pj_ok = (info >> 21) & 1u, pk_ok = (info >> 22) & 1u;
            uint32_t fresh = 0;                                       // bit m: neighbour m was not yet scheduled
            #pragma unroll
            for (int m = 1; m < 8; ++m) {                             // the atomics are independent: all in flight at once
                const bool a = m & 1, b = m & 2, cc = m & 4;
                if ((a && !pi_ok) || (b && !pj_ok) || (cc && !pk_ok)) continue;
                const int64_t d = c - (a ? si : 0) - (b ? sj : 0) - (cc ? sk : 0);
                if (use_hash) {
                    uint32_t h = ((uint32_t)d * 2654435761u) >> 21;                   // 11 bits
                    for (;;) {
                        const uint32_t old = atomicCAS(&sh.hset[h], HSET_EMPTY, (uint32_t)d);
                        if (old == HSET_EMPTY) { fresh |= 1u << m; break; }
                        if (old == (uint32_t)d) break;
                        h = (h + 1) & (HSET - 1);
                    }
                    continue;
                }
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
    RELAX_PHASE(2);
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

// ---- lookahead: one pass over the cells for up to eight consecutive sweeps (see LookState above) ---------------------
constexpr int LK_WARPS = 8;
constexpr int LK_THREADS = LK_WARPS * 32;
#ifndef SDFB_LOOK_QCAP
#define SDFB_LOOK_QCAP 512
#endif
#ifndef SDFB_LOOK_PCAP
#define SDFB_LOOK_PCAP 128
#endif
constexpr int LK_QCAP = SDFB_LOOK_QCAP;      // candidate queue entries per warp
constexpr int LK_PCAP = SDFB_LOOK_PCAP;      // pending voxels per warp
struct LookParams {
    Grid g;
    const uint64_t *cells;
    const TriRec *rec;
    int sweep_of[8];                         // direction q = s % 8 -> sweep index s of the window, -1 if the window has none
    uint32_t thr[8][8][8];                   // [q][class][m]: neighbour word >= thr <=> fresh for that sweep (0xffffffff: never)
    uint32_t tmin;                           // lowest threshold of the window: the cheap "nothing is fresh" test
    uint32_t othr[26];                       // class-0 threshold of each neighbour offset in the order of kLookOrder (standard windows)
    int standard;                            // the window starts at a multiple of 8: kLookOrder is its order of first examination
    int pass2;                               // ... and it is sweeps 8..15 with the thresholds of look_thr8 (compile-time filter)
    LookState *look;
    uint32_t *w_list;
    uint32_t cap_w;
    unsigned long long *changed;             // [1] += distance evaluations
};
struct LookPend { uint32_t c; float px, py, pz, phi; };
struct LookShared {
    uint32_t priv[LK_WARPS][26 * 32];        // the candidates of each lane's voxel (triangle | q << 27), entry e of lane l at e*32 + l
    uint32_t q_ent[LK_WARPS][LK_QCAP];       // triangle | q << 27
    uint16_t q_own[LK_WARPS][LK_QCAP];       // index into pend
    LookPend pend[LK_WARPS][LK_PCAP];
};

// directions as compile-time functions of q (SweepDir::of(q), cpu_lib/makelevelset3.cpp:245-248)
__host__ __device__ constexpr int look_di(int q) { return (q & 1) ? -1 : 1; }
__host__ __device__ constexpr int look_dj(int q) { return ((q & 1) ? -1 : 1) * ((((q % 8) >> 1) & 2) ? -1 : 1); }
__host__ __device__ constexpr int look_dk(int q) { return ((q & 1) ? -1 : 1) * ((((q % 8) >> 1) & 1) ? -1 : 1); }

__device__ __forceinline__ void look_mark(const LookParams &P, uint32_t c, int q)
{
    const unsigned idx = atomicAdd(&P.look->w_count[q], 1u);
    if (idx < P.cap_w) P.w_list[(size_t)q * P.cap_w + idx] = c;
    else P.look->dense_off = 1u;
}

// The candidates sweep direction Q would evaluate at this voxel: bit m = neighbour m (cpu_lib/makelevelset3.cpp:143-149).
// w[] = the {stamp|tri} words of the 3 x 3 x 3 neighbourhood, index (dk+1)*9 + (dj+1)*3 + (di+1) in ABSOLUTE offsets.
// UNIFORM = every lane of the warp is an interior voxel: class 0 thresholds, which are immediate operands then.
template <int Q, bool UNIFORM>
__device__ __forceinline__ uint32_t look_live(const LookParams &P, const uint32_t (&w)[27], uint32_t own, bool valid, int cls)
{
    constexpr int di = look_di(Q), dj = look_dj(Q), dk = look_dk(Q);
    uint32_t nb[7];
    #pragma unroll
    for (int m = 0; m < 7; ++m) {
        const int ci = (m == 0 || m == 2 || m == 4 || m == 6) ? 1 : 0, cj = (m == 1 || m == 2 || m == 5 || m == 6) ? 1 : 0, ck = (m >= 3) ? 1 : 0;
        nb[m] = w[(1 - dk * ck) * 9 + (1 - dj * cj) * 3 + (1 - di * ci)];
    }
    uint32_t live = 0;
    #pragma unroll
    for (int m = 0; m < 7; ++m) {
        const uint32_t x = nb[m];
        const uint32_t t = UNIFORM ? P.thr[Q][0][m] : P.thr[Q][cls][m];
        const bool keep = valid && ((x & TRI_MASK) != TRI_NONE) && (((x ^ own) & TRI_MASK) != 0) && (x >= t);
        live |= keep ? (1u << m) : 0u;
    }
    if (live) {                               // drop repeats of ANY earlier neighbour's triangle, as the sweeps do
        #pragma unroll
        for (int m = 1; m < 7; ++m) {
            bool dup = false;
            #pragma unroll
            for (int u = 0; u < m; ++u) dup = dup || (((nb[u] ^ nb[m]) & TRI_MASK) == 0);
            if (dup) live &= ~(1u << m);
        }
    }
    return live;
}

// Interior voxels of a window that starts at a multiple of 8: each of the 26 neighbour offsets can be fresh only in the
// FIRST sweep of the window that examines it (the later ones find it in their memo tables), so the filter walks the
// offsets once, in the order of that first examination: {di, dj, dk (absolute), q = direction of that sweep, m = the
// neighbour's number in that sweep}.  launch_look_scan checks the table against SweepDir::of.
struct LookOfs { int oi, oj, ok, q, m; };
__device__ constexpr LookOfs kLookOrder[26] = {
    {-1,0,0,0,0}, {0,-1,0,0,1}, {-1,-1,0,0,2}, {0,0,-1,0,3}, {-1,0,-1,0,4}, {0,-1,-1,0,5}, {-1,-1,-1,0,6},
    {1,0,0,1,0}, {0,1,0,1,1}, {1,1,0,1,2}, {0,0,1,1,3}, {1,0,1,1,4}, {0,1,1,1,5}, {1,1,1,1,6},
    {-1,0,1,2,4}, {0,-1,1,2,5}, {-1,-1,1,2,6}, {1,0,-1,3,4}, {0,1,-1,3,5}, {1,1,-1,3,6},
    {-1,1,0,4,2}, {-1,1,-1,4,6}, {1,-1,0,5,2}, {1,-1,1,5,6}, {-1,1,1,6,6}, {1,-1,-1,7,6}};
constexpr LookOfs kLookOrderHost[26] = {
    {-1,0,0,0,0}, {0,-1,0,0,1}, {-1,-1,0,0,2}, {0,0,-1,0,3}, {-1,0,-1,0,4}, {0,-1,-1,0,5}, {-1,-1,-1,0,6},
    {1,0,0,1,0}, {0,1,0,1,1}, {1,1,0,1,2}, {0,0,1,1,3}, {1,0,1,1,4}, {0,1,1,1,5}, {1,1,1,1,6},
    {-1,0,1,2,4}, {0,-1,1,2,5}, {-1,-1,1,2,6}, {1,0,-1,3,4}, {0,1,-1,3,5}, {1,1,-1,3,6},
    {-1,1,0,4,2}, {-1,1,-1,4,6}, {1,-1,0,5,2}, {1,-1,1,5,6}, {-1,1,1,6,6}, {1,-1,-1,7,6}};
__host__ __device__ constexpr int look_widx(int oi, int oj, int ok) { return (ok + 1) * 9 + (oj + 1) * 3 + (oi + 1); }
// Which earlier offsets (in window order) a candidate is compared with before it is kept: a triangle that an earlier
// offset already names is either measured there (and if it wins, that earlier sweep changes the voxel, which puts it on
// c_list, so every later sweep looks at the voxel anyway), or it is the voxel's own, or the memo says it lost before.
// SDFB_LOOK_DEDUPE = 0: the sweeps' own rule (the earlier neighbours of the same sweep); d > 0: every earlier offset within
// Manhattan distance d (neighbouring voxels share triangles; 1 = face-adjacent offsets, 48 pairs in all).
// Measured at C2 (profiles/r2_lookahead.txt): 0 -> 1.71 evaluations per voxel in the second pass, 1 -> 1.97, 2 -> 1.68 (more compares).
#ifndef SDFB_LOOK_DEDUPE
#define SDFB_LOOK_DEDUPE 0
#endif
#if SDFB_LOOK_DEDUPE > 0
__host__ __device__ constexpr int look_abs(int v) { return v < 0 ? -v : v; }
#endif
__device__ constexpr bool look_partner(int n, int u)
{
    if (u >= n) return false;
    const LookOfs a = kLookOrder[n], b = kLookOrder[u];
#if SDFB_LOOK_DEDUPE > 0
    return look_abs(a.oi - b.oi) + look_abs(a.oj - b.oj) + look_abs(a.ok - b.ok) <= SDFB_LOOK_DEDUPE;
#else
    // neighbour u' < m of sweep a.q sits at offset -d * c(u')
    for (int up = 0; up < a.m; ++up) {
        const int oi = -look_di(a.q) * ((up == 0 || up == 2 || up == 4 || up == 6) ? 1 : 0),
                  oj = -look_dj(a.q) * ((up == 1 || up == 2 || up == 5 || up == 6) ? 1 : 0), ok = -look_dk(a.q) * ((up >= 3) ? 1 : 0);
        if (b.oi == oi && b.oj == oj && b.ok == ok) return true;
    }
    return false;
#endif
}

template <int N, int U> struct LookDup {      // does any partner offset u <= U of offset N hold triangle x?
    static __device__ __forceinline__ bool any(const uint32_t (&w)[27], uint32_t x)
    {
        bool d = LookDup<N, U - 1>::any(w, x);
        if (look_partner(N, U)) { constexpr LookOfs b = kLookOrder[U]; d = d || (((w[look_widx(b.oi, b.oj, b.ok)] ^ x) & TRI_MASK) == 0); }
        return d;
    }
};
template <int N> struct LookDup<N, -1> { static __device__ __forceinline__ bool any(const uint32_t (&)[27], uint32_t) { return false; } };

// The window that matters is the reference's second pass, sweeps 8..15 after sweeps 0..7: its thresholds are a function of
// the direction table alone (memo_last_table), so they can be compile-time constants -- and the offsets that sweep 7 was the
// last to examine (stamp threshold 9) cannot be fresh at all, because no cell carries a stamp above 8 before sweep 8.
// look_thr8(q, m) = lowest stamp that makes neighbour m of sweep 8 + q fresh for an interior voxel.
__host__ __device__ constexpr int look_thr8(int q, int m)
{
    const bool ci = (m == 0 || m == 2 || m == 4 || m == 6), cj = (m == 1 || m == 2 || m == 5 || m == 6), ck = (m >= 3);
    for (int e = 8 + q - 1; e >= 0; --e) {
        const int qe = e & 7;
        if ((!ci || look_di(qe) == look_di(q)) && (!cj || look_dj(qe) == look_dj(q)) && (!ck || look_dk(qe) == look_dk(q))) return e + 2;
    }
    return 1;
}

template <int N, bool PASS2> struct LookStep {
    // Offsets 0..N in window order: a neighbour that names another triangle than the voxel's, whose cell is newer than the
    // memo entry of the sweep that examines it first and whose triangle no partner offset repeats goes to the lane's private
    // list (entry = triangle | q << 27).  A neighbour without a triangle carries stamp 0 and fails every threshold of a second
    // pass (launch_look_scan checks that they are all >= 1 << 27).  PASS2: the window is sweeps 8..15 (see look_thr8).
    static __device__ __forceinline__ void filter(const LookParams &P, const uint32_t (&w)[27], uint32_t own, uint32_t *priv, int &cnt)
    {
        LookStep<N - 1, PASS2>::filter(P, w, own, priv, cnt);
        constexpr LookOfs o = kLookOrder[N];
        constexpr int t8 = look_thr8(o.q, o.m);
        if (PASS2 && t8 > 8) return;                                  // compile-time: this offset is never fresh in that window
        const uint32_t x = w[look_widx(o.oi, o.oj, o.ok)];
        const uint32_t thr = PASS2 ? (uint32_t)t8 << 27 : P.othr[N];
        if ((((x ^ own) & TRI_MASK) != 0) && (x >= thr) && !LookDup<N, N - 1>::any(w, x)) {
            priv[cnt * 32] = (x & TRI_MASK) | ((uint32_t)o.q << 27);
            ++cnt;
        }
    }
};
template <bool PASS2> struct LookStep<-1, PASS2> {
    static __device__ __forceinline__ void filter(const LookParams &, const uint32_t (&)[27], uint32_t, uint32_t *, int &) {}
};

// generic path: the live neighbours of direction Q go to the lane's private list, like the table-driven path's
// (the list holds 26 entries -- every neighbour offset once; a voxel of an odd window that has more is put on the sweep's
// list unevaluated instead)
template <int Q>
__device__ __forceinline__ void look_enqueue(const LookParams &P, uint32_t c, const uint32_t (&w)[27], uint32_t live, uint32_t *priv, int &cnt)
{
    constexpr int di = look_di(Q), dj = look_dj(Q), dk = look_dk(Q);
    if (!live) return;
    #pragma unroll
    for (int m = 0; m < 7; ++m) if ((live >> m) & 1u) {
        const int ci = (m == 0 || m == 2 || m == 4 || m == 6) ? 1 : 0, cj = (m == 1 || m == 2 || m == 5 || m == 6) ? 1 : 0, ck = (m >= 3) ? 1 : 0;
        if (cnt < 26) { priv[cnt * 32] = (w[(1 - dk * ck) * 9 + (1 - dj * cj) * 3 + (1 - di * ci)] & TRI_MASK) | ((uint32_t)Q << 27); ++cnt; }
        else { look_mark(P, c, Q); break; }
    }
}

// Evaluate the warp's queue: an entry that beats its voxel's distance puts the voxel on the list of the entry's sweep.
__device__ __forceinline__ void look_flush(const LookParams &P, LookShared &sh, int warp, int lane, int &nq, int &np, unsigned &evals)
{
    __syncwarp();
    for (int e = lane; e < nq; e += 32) {
        const uint32_t ent = sh.q_ent[warp][e];
        const LookPend &pe = sh.pend[warp][sh.q_own[warp][e]];
        const TriRec *tr = &P.rec[ent & TRI_MASK];
        const float4 p = __ldg(&tr->p), qq = __ldg(&tr->q), r = __ldg(&tr->r);
        const float d = ptd_rec(F3{pe.px, pe.py, pe.pz}, p, qq, r);
        ++evals;
        if (d < pe.phi) look_mark(P, pe.c, (int)(ent >> 27));
    }
    __syncwarp();
    nq = 0; np = 0;
}

#ifndef SDFB_LOOK_MINB
#define SDFB_LOOK_MINB 3
#endif
__global__ void __launch_bounds__(LK_THREADS, SDFB_LOOK_MINB) k_look_scan(const __grid_constant__ LookParams P)
{
    extern __shared__ __align__(16) unsigned char look_smem[];
    LookShared &sh = *reinterpret_cast<LookShared *>(look_smem);
    const Grid &g = P.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int jblocks = (g.nj + LK_WARPS - 1) / LK_WARPS;
    const int64_t nitems = (int64_t)jblocks * g.nk;
    const int64_t plane = g.plane();
    uint32_t *const priv = &sh.priv[warp][lane];
    unsigned evals = 0;
    int nq = 0, np = 0;
    // the first toucher of plane k+1 is the item of plane k: ask L2 for that row one item ahead
    auto prefetch_row = [&](int64_t item) {
        if (item >= nitems) return;
        const int pk = (int)(item / jblocks) + 1, pj = (int)(item % jblocks) * LK_WARPS + warp;
        if (pj > g.nj - 1 || pk > g.nk - 1) return;
        const uintptr_t beg = reinterpret_cast<uintptr_t>(P.cells + g.cidx(0, pj, pk)) & ~(uintptr_t)15;
        const uint32_t bytes = ((uint32_t)g.ni * 8u + 16u) & ~15u;
        for (uint32_t o = (uint32_t)lane * 2048u; o < bytes; o += 32u * 2048u) {
            const uint32_t len = min(2048u, bytes - o);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(beg + o), "r"(len) : "memory");
        }
    };
    prefetch_row(blockIdx.x);
    for (int64_t it = blockIdx.x; it < nitems; it += gridDim.x) {
        prefetch_row(it + gridDim.x);
        if (__ldcg(&P.look->dense_off)) { nq = np = 0; break; }        // a list overflowed: the sweeps will scan for themselves
        const int k = (int)(it / jblocks), j = (int)(it % jblocks) * LK_WARPS + warp;
        if (j > g.nj - 1) continue;                                   // warp-uniform
        const int64_t row = g.cidx(0, j, k);
        const uint32_t *base = reinterpret_cast<const uint32_t *>(P.cells + row);      // low words: {stamp | tri}
        const bool row_interior = j >= 1 && j <= g.nj - 2 && k >= 1 && k <= g.nk - 2;
        const float py = lattice(j, g.dx, g.oy), pz = lattice(k, g.dx, g.oz);
        // interior voxels of the row (1 <= i <= ni-2) in chunks of 32, then one chunk with the two face voxels i = 0 and
        // i = ni-1: whole warps of interior voxels of an interior row take the table-driven path
        const int nint = (g.ni - 2 + 31) / 32;
        for (int ch = 0; ch <= nint; ++ch) {
            const bool edge = ch == nint;
            const int i = edge ? (lane == 0 ? 0 : g.ni - 1) : 1 + ch * 32 + lane;
            const bool inb = edge ? lane < 2 : i <= g.ni - 2;
            const bool fast = P.standard && row_interior && !edge;                // warp-uniform
            uint32_t w[27], live[8];
            int ncand = 0;
            if (fast) {
                // all 27 cells exist (lanes past the end of the row read the next row's first cells and are masked below)
                #pragma unroll
                for (int ok = -1; ok <= 1; ++ok) {
                    #pragma unroll
                    for (int oj = -1; oj <= 1; ++oj) {
                        const uint32_t *pp = base + 2 * ((int64_t)oj * g.ni + (int64_t)ok * plane + i);
                        #pragma unroll
                        for (int oi = -1; oi <= 1; ++oi) w[(ok + 1) * 9 + (oj + 1) * 3 + (oi + 1)] = __ldg(pp + 2 * oi);
                    }
                }
                uint32_t mx = 0;
                #pragma unroll
                for (int t = 0; t < 27; ++t) if (t != 13) mx = max(mx, w[t]);
                if (__any_sync(0xffffffffu, inb && mx >= P.tmin)) {
                    if (P.pass2) LookStep<25, true>::filter(P, w, w[13], priv, ncand);
                    else LookStep<25, false>::filter(P, w, w[13], priv, ncand);
                    if (!inb) ncand = 0;
                }
            } else {
                #pragma unroll
                for (int ok = -1; ok <= 1; ++ok) {
                    #pragma unroll
                    for (int oj = -1; oj <= 1; ++oj) {
                        const bool rok = (unsigned)(j + oj) < (unsigned)g.nj && (unsigned)(k + ok) < (unsigned)g.nk;   // warp-uniform
                        const uint32_t *rp = base + 2 * ((int64_t)oj * g.ni + (int64_t)ok * plane);
                        #pragma unroll
                        for (int oi = -1; oi <= 1; ++oi) {
                            const int ii = i + oi;
                            w[(ok + 1) * 9 + (oj + 1) * 3 + (oi + 1)] = (rok && inb && (unsigned)ii < (unsigned)g.ni) ? __ldg(rp + 2 * ii) : TRI_NONE;
                        }
                    }
                }
                const uint32_t own = w[13];
                #pragma unroll
                for (int q = 0; q < 8; ++q) live[q] = 0;
                // a voxel on the face a sweep starts from is not updated by it; one on the opposite face has its own memo class
                const bool i_lo = i == 0, i_hi = i == g.ni - 1, j_lo = j == 0, j_hi = j == g.nj - 1, k_lo = k == 0, k_hi = k == g.nk - 1;
#define SDFB_LOOK_Q(Q) if (P.sweep_of[Q] >= 0) { \
                    const bool st = (look_di(Q) > 0 ? i_lo : i_hi) || (look_dj(Q) > 0 ? j_lo : j_hi) || (look_dk(Q) > 0 ? k_lo : k_hi); \
                    const int cls = ((look_di(Q) > 0 ? i_hi : i_lo) ? 1 : 0) | ((look_dj(Q) > 0 ? j_hi : j_lo) ? 2 : 0) | ((look_dk(Q) > 0 ? k_hi : k_lo) ? 4 : 0); \
                    live[Q] = look_live<Q, false>(P, w, own, inb && !st, cls); }
                SDFB_LOOK_Q(0) SDFB_LOOK_Q(1) SDFB_LOOK_Q(2) SDFB_LOOK_Q(3) SDFB_LOOK_Q(4) SDFB_LOOK_Q(5) SDFB_LOOK_Q(6) SDFB_LOOK_Q(7)
#undef SDFB_LOOK_Q
                const uint32_t cc = (uint32_t)(row + i);
                look_enqueue<0>(P, cc, w, live[0], priv, ncand); look_enqueue<1>(P, cc, w, live[1], priv, ncand);
                look_enqueue<2>(P, cc, w, live[2], priv, ncand); look_enqueue<3>(P, cc, w, live[3], priv, ncand);
                look_enqueue<4>(P, cc, w, live[4], priv, ncand); look_enqueue<5>(P, cc, w, live[5], priv, ncand);
                look_enqueue<6>(P, cc, w, live[6], priv, ncand); look_enqueue<7>(P, cc, w, live[7], priv, ncand);
            }
            const uint32_t bp = __ballot_sync(0xffffffffu, ncand > 0);
            if (!bp) continue;
            int incl = ncand;                                         // inclusive warp scan of the candidate counts
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            const int64_t c = row + i;
            if (nq + total > LK_QCAP || np + 32 > LK_PCAP) look_flush(P, sh, warp, lane, nq, np, evals);
            // Enqueue the chunk, lane range by lane range if its candidates do not fit the (now empty) queue at once: a
            // voxel has at most 56, so every range makes progress.
            int start = 0, done = 0;                                  // first lane not yet enqueued, candidates before it
            for (;;) {
                const uint32_t fits = __ballot_sync(0xffffffffu, lane >= start && incl - done <= LK_QCAP - nq);
                const int end = start + __popc(fits);                 // the lanes that fit are a prefix of [start, 32)
                const bool mine = lane >= start && lane < end;
                const uint32_t range = (end >= 32 ? 0xffffffffu : (1u << end) - 1u) & ~((1u << start) - 1u);
                if (mine && ncand > 0) {
                    const int pidx = np + __popc(bp & range & ((1u << lane) - 1u));
                    LookPend &pe = sh.pend[warp][pidx];
                    pe.c = (uint32_t)c;
                    pe.px = lattice(i, g.dx, g.ox); pe.py = py; pe.pz = pz;
                    pe.phi = __uint_as_float(__ldg(base + 2 * i + 1));    // high word: the voxel's distance
                    int wq = nq + incl - ncand - done;
                    for (int e = 0; e < ncand; ++e, ++wq) { sh.q_ent[warp][wq] = priv[e * 32]; sh.q_own[warp][wq] = (uint16_t)pidx; }
                }
                const int upto = __shfl_sync(0xffffffffu, incl, end - 1);   // candidates of lanes [0, end)
                nq += upto - done;
                np += __popc(bp & range);
                if (end >= 32) break;
                look_flush(P, sh, warp, lane, nq, np, evals);
                start = end; done = upto;
            }
        }
    }
    look_flush(P, sh, warp, lane, nq, np, evals);
    for (int o = 16; o > 0; o >>= 1) evals += __shfl_down_sync(0xffffffffu, evals, o);
    if (lane == 0 && evals) atomicAdd(P.changed + 1, (unsigned long long)evals);
}

// Round-0 work of sweep P.stamp - 1 from the lookahead lists: W of the sweep's direction, and every cell changed since the
// scan together with its seven downstream neighbours (those the sweep updates).
__global__ void __launch_bounds__(256) k_look_mark(RelaxParams P)
{
    // the decision "this sweep starts from the window's lists" is taken HERE, once, and handed to the rounds kernel in
    // count[7]: dense_off itself may be raised while that kernel runs (c_list overflow), and its CTAs must all agree
    const bool off = __ldcg(&P.look->dense_off) != 0u;
    if (blockIdx.x == 0 && threadIdx.x == 0) P.count[7] = off ? 0u : 1u;
    if (off) return;
    const Grid &g = P.g;
    const int q = (int)((P.stamp - 1u) & 7u);
    const unsigned nW = min(__ldcg(&P.look->w_count[q]), P.cap_w), nC = min(__ldcg(&P.look->c_count), P.cap_c);
    const uint64_t total = (uint64_t)nW + 8ull * nC;
    const uint32_t plane32 = (uint32_t)g.plane();
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        int64_t d;
        if (t < nW) {
            d = (int64_t)__ldcg(&P.w_list[(size_t)q * P.cap_w + t]);
        } else {
            const uint64_t e = (t - nW) >> 3;
            const int m = (int)((t - nW) & 7u);
            const uint32_t c = __ldcg(&P.c_list[e]);
            const uint32_t p = c / plane32, rem = c - p * plane32;
            const int j = (int)(rem / (uint32_t)g.ni), i = (int)(rem - (uint32_t)j * (uint32_t)g.ni), k = (int)p - 1 + g.k_lo;
            const int ri = (P.sd.di > 0 ? i : g.ni - 1 - i) + (m & 1), rj = (P.sd.dj > 0 ? j : g.nj - 1 - j) + ((m >> 1) & 1),
                      rk = P.sd.rel_k(k, g) + ((m >> 2) & 1);
            if (ri < 1 || ri > g.ni - 1 || rj < 1 || rj > g.nj - 1 || rk < P.rk_first || rk > P.rk_last) continue;
            d = g.cidx(P.sd.abs_i(ri, g), P.sd.abs_j(rj, g), P.sd.abs_k(rk, g));
        }
        const uint32_t bit = 1u << (d & 31);
        if (!(atomicOr(&P.bitmap[0][d >> 5], bit) & bit)) {           // not yet on the list of round 0
            const unsigned idx = atomicAdd(&P.count[0], 1u);
            if (idx < P.list_cap) P.list[0][idx] = (uint32_t)d;       // (beyond that the round walks the bitmap)
        }
    }
}

// What a lookahead window changed, as patches for a signed-distance output that was produced (and is being copied to the
// host) from the cells as they were BEFORE the window: every entry of c_list -> {index in the output layout, final value}.
// The sign of a voxel depends on the crossing counts only, so it is taken from the early output; the magnitude is the
// cell's.  head[0] = entries written (0 if they do not fit `cap`), head[1] = c_count, head[2] = dense_off: the caller
// may only use the patches when head[2] == 0 and head[1] <= cap.
__global__ void __launch_bounds__(256) k_look_patches(const uint64_t *__restrict__ cells, const float *__restrict__ phi_early, Grid g,
                                                      int kfastest, const LookState *__restrict__ look, const uint32_t *__restrict__ c_list,
                                                      uint32_t cap_c, uint32_t *__restrict__ patch_idx, float *__restrict__ patch_val,
                                                      uint32_t cap, unsigned int *__restrict__ head)
{
    const unsigned n_all = look->c_count, off = look->dense_off;
    const bool usable = off == 0u && n_all <= cap && n_all <= cap_c;
    if (blockIdx.x == 0 && threadIdx.x == 0) { head[0] = usable ? n_all : 0u; head[1] = n_all; head[2] = off; }
    if (!usable) return;
    const uint32_t plane32 = (uint32_t)g.plane();
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n_all; e += gridDim.x * blockDim.x) {
        const uint32_t c = c_list[e];
        const uint32_t p = c / plane32, rem = c - p * plane32;
        const uint32_t j = rem / (uint32_t)g.ni, i = rem - j * (uint32_t)g.ni, k = p - 1u;       // whole-grid plans: k_lo = 0
        const int64_t v = (int64_t)c - g.plane();
        const float mag = cell_phi(cells[c]);
        const bool inside = (__float_as_uint(phi_early[v]) >> 31) != 0u;
        patch_idx[e] = kfastest ? (uint32_t)(((int64_t)i * g.nj + j) * (int64_t)g.nk + k) : (uint32_t)v;
        patch_val[e] = inside ? -mag : mag;
    }
}

// Grid-wide barrier for the co-resident (cooperatively launched) CTAs: one arrival counter that only grows;
// `target` is the value it reaches when every CTA has arrived at this barrier.  Several times cheaper than
// cooperative_groups' grid.sync() here, and the rounds are all latency.
__device__ __forceinline__ void grid_barrier(unsigned int *ctr, unsigned int &target, unsigned int nctas)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        target += nctas;
        __threadfence();
        atomicAdd(ctr, 1u);
        while (*reinterpret_cast<volatile unsigned int *>(ctr) < target) { }
        __threadfence();
    }
    __syncthreads();
}

// ---- the kernel: round 0 over all voxels, then the rounds ---------------------------------------------------------------------------------------
constexpr int MAX_GRID_ROUNDS = 256;         // more grid-wide rounds than this cost as much as a column sweep
// 256 = HASH_MAX: every single-CTA round is one batch per warp and de-duplicates in shared memory (512: second pass of C2
// 12.9 ms, 256: 12.3, 128: 12.8, 64: 13.1 -- profiles/r2_lookahead.txt)
#ifndef SDFB_RELAX_SOLO_MAX
#define SDFB_RELAX_SOLO_MAX 256
#endif
constexpr unsigned SOLO_MAX = SDFB_RELAX_SOLO_MAX;   // lists this short are finished by one CTA (a CTA barrier per round
                                             // instead of a grid barrier)
// In between, a TEAM of the first few CTAs carries on: a list of a few thousand entries is one batch for their warps, and a
// barrier among 16 CTAs costs a fraction of one among all 444 (the rounds are pure latency: ~80 of them per sweep).
#ifndef SDFB_RELAX_TEAM
#define SDFB_RELAX_TEAM 16
#endif
constexpr unsigned TEAM_CTAS = SDFB_RELAX_TEAM;
constexpr unsigned TEAM_MAX = TEAM_CTAS * RX_WARPS * 32;    // entries the team handles in one batch

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
    // round 0 from a bitmap: the lean scan kernel filled it, or k_look_mark did (lookahead window still valid)
    const bool look_on = P.look && __ldcg(&P.count[7]) != 0u;        // set by k_look_mark, constant during this launch
    const bool scan_mode = P.scan_mode || look_on;
    if (!scan_mode) {
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
        grid_barrier(&P.count[4], bar_target, gridDim.x);
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
    int r = scan_mode ? 0 : 1;               // with a scan kernel, round 0 runs here from the bitmap it filled
    bool solo = false, team = false;
    unsigned int team_target = 0;
    unsigned long long work = 0;             // list entries so far (the same number in every CTA)
    for (;; ++r) {
        const int par = r & 1;
        unsigned n = solo ? *reinterpret_cast<volatile unsigned int *>(&sh.scount[r % 3])
                          : *reinterpret_cast<volatile unsigned int *>(&P.count[r % 3]);
        if (n == 0) break;
        RELAX_PHASE(3);
        work += n;
        // (not once the team or one CTA has taken over: the CTAs that left have already reported their change counts, and a
        // list of a few thousand entries is no longer the heavy case)
        if (!solo && !team && r >= 1 && (work > P.heavy_limit || r > MAX_GRID_ROUNDS)) {                // uniform over the grid
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
            if (blockIdx.x == 0 && tid == 0) {
                P.count[5] = 1u;
                if (P.look) P.look->dense_off = 1u;                   // the column schedule will not record what it changes
            }
            net_changed = 0;
            break;
        }
        if (P.debug && blockIdx.x == 0 && tid == 0 && r < 250) {
            unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            P.debug[4 + 2 * r] = n; P.debug[5 + 2 * r] = t - t_start;
        }
        if (!solo && !team && r > 0 && n <= TEAM_MAX && gridDim.x > TEAM_CTAS) {     // uniform over the grid
            team = true;
            if (blockIdx.x >= TEAM_CTAS) break;
        }
        if (!solo && r > 0 && n <= SOLO_MAX) {                        // uniform over the grid / the team
            solo = true;
            if (blockIdx.x != 0) break;
            for (unsigned t = tid; t < n; t += RX_THREADS) sh.slist[par][t] = __ldcg(&P.list[par][t]);     // import the list
            if (tid == 0) sh.scount[(r + 1) % 3] = 0;
#ifdef SDFB_RELAX_PHASES
            if (tid == 0) sh.tprev = clock64();
#endif
            __syncthreads();
        }
        if (solo) { if (tid == 0) sh.scount[(r + 2) % 3] = 0; }
        else if (blockIdx.x == 0 && tid == 0) P.count[(r + 2) % 3] = 0;
        unsigned int *const push_count = solo ? &sh.scount[(r + 1) % 3] : &P.count[(r + 1) % 3];
        // (only when every push is sure to fit the next list: a list that overflows is walked from the global bitmap)
        const bool use_hash = solo && n <= HASH_MAX && 7u * n <= (unsigned)SOLO_CAP + P.list_cap;      // uniform over the CTA
        if (use_hash) {
            for (int t = tid; t < HSET; t += RX_THREADS) sh.hset[t] = HSET_EMPTY;
            __syncthreads();
        }
        // the bitmap is the work list when a list overflowed
        // (round 0 after the lean scan kernel walks the bitmap it filled; k_look_mark leaves a list like any other round's)
        const bool use_bitmap = (r == 0 && !look_on) || n > P.list_cap + (solo ? (unsigned)SOLO_CAP : 0u);
        const int64_t w = solo ? warp : gwarp, nw = solo ? RX_WARPS : (team ? (int64_t)TEAM_CTAS * RX_WARPS : nwarps);
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
#ifdef SDFB_RELAX_PHASES
                asm volatile("" ::"l"(base64), "r"(wd.nb[0] ^ wd.nb[1] ^ wd.nb[2] ^ wd.nb[3] ^ wd.nb[4] ^ wd.nb[5] ^ wd.nb[6]));
                RELAX_PHASE(4);
#endif
                relax_filter(P, sh, warp, lane, valid, c, ri, rj, rk, wd.own, base64, wd.nb, was_changed, nq, np);
                if (nq > QCAP_B - 7 * 32 || np > PCAP_B - 32) relax_flush(P, sh, warp, lane, nq, np, par ^ 1, push_count, solo, net_changed, evals, use_hash);
            }
        }
        relax_flush(P, sh, warp, lane, nq, np, par ^ 1, push_count, solo, net_changed, evals, use_hash);
        if (solo) __syncthreads();                                    // orders the CTA's writes (global and shared) and reads
        else if (team) grid_barrier(&P.count[6], team_target, TEAM_CTAS);
        else grid_barrier(&P.count[4], bar_target, gridDim.x);
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

// scratch the schedule needs for a slab: three counters, oldbuf (8 B per cell), two lists, two bitmaps, and the lookahead
// window's state and lists
bool sweep_relax_supported(const Grid &g) { return g.cell_count() < ((int64_t)1 << 32); }
uint32_t sweep_relax_list_cap(const Grid &g)
{
    const int64_t cap = g.cell_count() < ((int64_t)16 << 20) ? g.cell_count() : ((int64_t)16 << 20);
    return (uint32_t)cap;
}
namespace {
struct RelaxLayout { size_t count, oldbuf, list0, list1, bitmap0, bitmap1, look, c_list, w_list, total; uint32_t list_cap, cap_c, cap_w; };
RelaxLayout relax_layout(const Grid &g)
{
    RelaxLayout L{};
    const size_t cells = (size_t)g.cell_count(), words = (cells + 31) / 32 + 1;
    L.list_cap = sweep_relax_list_cap(g);
    // lookahead lists: a light second pass changes ~0.1 % of the cells in all and marks ~0.02 .. 0.1 % of them per sweep;
    // anything that does not fit a sixteenth (changed cells) / a sixty-fourth (marks of one sweep) is the heavy case
    auto clamp = [](size_t v, size_t lo, size_t hi) { return v < lo ? lo : (v > hi ? hi : v); };
    L.cap_c = (uint32_t)clamp(cells / 16, cells < 4096 ? cells : 4096, (size_t)16 << 20);
    L.cap_w = (uint32_t)clamp(cells / 64, cells < 4096 ? cells : 4096, (size_t)4 << 20);
    size_t o = 0;
    L.count = o; o += 64;
    L.oldbuf = o; o += cells * 8;
    L.list0 = o; o += (size_t)L.list_cap * 4;
    L.list1 = o; o += (size_t)L.list_cap * 4;
    L.bitmap0 = o; o += words * 4;
    L.bitmap1 = o; o += words * 4;
    L.look = o; o += sizeof(LookState);
    L.c_list = o; o += (size_t)L.cap_c * 4;
    L.w_list = o; o += (size_t)L.cap_w * 4 * 8;
    L.total = o;
    return L;
}
}  // namespace
size_t sweep_relax_scratch_bytes(const Grid &g) { return relax_layout(g).total; }

// word that is non-zero after a launch that gave its sweep back (launch_sweep_columns' run_if)
const unsigned int *sweep_relax_fallback_flag(const void *scratch) { return static_cast<const unsigned int *>(scratch) + 5; }

// `scratch` must be zero-initialised once after allocation (bitmaps and counters return to zero after every sweep).
int launch_sweep_relax(uint64_t *cells, const TriRec *rec, const Grid &g, int sweep_index,
                       unsigned long long *changed, void *scratch, cudaStream_t st, const Tuning &tun, int max_ctas, bool look)
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
    const size_t ncells = (size_t)g.cell_count();
    const RelaxLayout L = relax_layout(g);
    char *s = static_cast<char *>(scratch);
    P.count = reinterpret_cast<unsigned int *>(s + L.count);
    P.oldbuf = reinterpret_cast<uint64_t *>(s + L.oldbuf);
    P.list[0] = reinterpret_cast<uint32_t *>(s + L.list0);
    P.list[1] = reinterpret_cast<uint32_t *>(s + L.list1);
    P.bitmap[0] = reinterpret_cast<uint32_t *>(s + L.bitmap0);
    P.bitmap[1] = reinterpret_cast<uint32_t *>(s + L.bitmap1);
    if (look) {
        P.look = reinterpret_cast<LookState *>(s + L.look);
        P.c_list = reinterpret_cast<uint32_t *>(s + L.c_list); P.cap_c = L.cap_c;
        P.w_list = reinterpret_cast<uint32_t *>(s + L.w_list); P.cap_w = L.cap_w;
        if (tun.look_cap > 0) {                                       // tests: force the overflow paths (same caps as launch_look_scan:
            P.cap_c = min(P.cap_c, (uint32_t)tun.look_cap);           // cap_w is also the row stride of w_list)
            P.cap_w = min(P.cap_w, (uint32_t)tun.look_cap);
        }
    }
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
    P.scan_mode = (!look && sweep_index >= tun.relax_scan_from) ? 1 : 0;
    int launches = 1;
    if (look) {
        // round 0 from the lookahead lists (the kernel does its own dense pass if the window was switched off on the device)
        k_look_mark<<<sms * 2, 256, 0, st>>>(P);
        ++launches;
    }
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
#ifdef SDFB_RELAX_PHASES
        fprintf(stderr, "[relax] sweep %2d single-CTA rounds, cycles of warp 0: loads %llu filter %llu eval %llu replay+push %llu barrier %llu\n", sweep_index,
                h[304], h[300], h[301], h[302], h[303]);
#endif
        fprintf(stderr, "[relax] sweep %2d: round 0 %.3f ms, total %.3f ms, first list %llu, rounds %llu, grid %d x %d\n", sweep_index,
                h[0] * 1e-6, h[1] * 1e-6, h[2], h[3], grid, RX_THREADS);
    }
    return launches;
}


// Lookahead over sweeps s_lo .. s_hi-1 (at most eight, so that every direction occurs once): zeroes the window's state and
// runs k_look_scan on the cells as they are now.  The sweeps of the window must follow in order, each through
// launch_sweep_relax(..., look = true); anything else invalidates the window (the caller then passes look = false).
int launch_look_scan(const uint64_t *cells, const TriRec *rec, const Grid &g, int s_lo, int s_hi,
                     unsigned long long *changed, void *scratch, cudaStream_t st, const Tuning &tun, int max_ctas)
{
    if (s_hi - s_lo < 1 || s_hi - s_lo > 8 || g.ni < 2 || g.nj < 2 || g.nk < 2 || g.k_lo != 0 || g.k_hi != g.nk) return 0;
    const RelaxLayout L = relax_layout(g);
    char *sc = static_cast<char *>(scratch);
    LookParams P{};
    P.g = g; P.cells = cells; P.rec = rec; P.changed = changed;
    P.look = reinterpret_cast<LookState *>(sc + L.look);
    P.w_list = reinterpret_cast<uint32_t *>(sc + L.w_list); P.cap_w = L.cap_w;
    if (tun.look_cap > 0) P.cap_w = min(P.cap_w, (uint32_t)tun.look_cap);
    P.tmin = 0xffffffffu;
    for (int q = 0; q < 8; ++q) {
        P.sweep_of[q] = -1;
        for (int c = 0; c < 8; ++c) for (int m = 0; m < 8; ++m) P.thr[q][c][m] = 0xffffffffu;
    }
    for (int s = s_lo; s < s_hi; ++s) {
        const int q = s & 7;
        const SweepDir sd = SweepDir::of(s);
        uint8_t last[8][8];
        memo_last_table(s, sd, last);
        P.sweep_of[q] = s;
        for (int c = 0; c < 8; ++c) for (int m = 0; m < 7; ++m) {
            const uint32_t l = last[c][m];
            P.thr[q][c][m] = l ? (l + 1u) << 27 : 0u;
            if (P.thr[q][c][m] < P.tmin) P.tmin = P.thr[q][c][m];
        }
    }
    // interior voxels of a window that starts at a multiple of 8 walk the 26 offsets once (kLookOrder); the table must be
    // the order in which SweepDir::of's directions first examine each offset
    P.standard = (s_lo & 7) == 0 ? 1 : 0;
    {
        bool seen[27] = {false};
        int n = 0;
        for (int q = 0; q < 8 && P.standard; ++q) {
            const SweepDir sd = SweepDir::of(q);
            for (int m = 0; m < 7; ++m) {
                const int oi = -sd.di * ((m == 0 || m == 2 || m == 4 || m == 6) ? 1 : 0), oj = -sd.dj * ((m == 1 || m == 2 || m == 5 || m == 6) ? 1 : 0),
                          ok = -sd.dk * (m >= 3 ? 1 : 0);
                if (seen[look_widx(oi, oj, ok)]) continue;
                seen[look_widx(oi, oj, ok)] = true;
                if (n >= 26) { P.standard = 0; break; }
                const LookOfs &o = kLookOrderHost[n];
                if (o.oi != oi || o.oj != oj || o.ok != ok || o.q != q || o.m != m) { P.standard = 0; break; }
                P.othr[n] = P.sweep_of[q] >= 0 ? P.thr[q][0][m] : 0xffffffffu;
                ++n;
            }
        }
        if (n != 26) P.standard = 0;
        // the table-driven filter leaves out the "names a triangle" test: a cell without one has stamp 0
        for (int t = 0; t < 26 && P.standard; ++t) if (P.othr[t] < (1u << 27)) P.standard = 0;
        // the reference's second pass as a whole: thresholds as compile-time constants (checked against the table just built)
        P.pass2 = (P.standard && s_lo == 8 && s_hi == 16) ? 1 : 0;
        for (int t = 0; t < 26 && P.pass2; ++t)
            if (P.othr[t] != (uint32_t)look_thr8(kLookOrderHost[t].q, kLookOrderHost[t].m) << 27) P.pass2 = 0;
    }
    cudaMemsetAsync(P.look, 0, sizeof(LookState), st);
    int dev = 0, sms = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = sizeof(LookShared);
    static int occ_cached[64] = {0};          // per device, as in launch_sweep_relax
    if (dev < 0 || dev >= 64 || !occ_cached[dev]) {
        cudaFuncSetAttribute(k_look_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_look_scan, LK_THREADS, smem);
        if (occ < 1) occ = 1;
        if (dev >= 0 && dev < 64) occ_cached[dev] = occ;
    } else occ = occ_cached[dev];
    int grid = sms * occ;
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    const int64_t nitems = (int64_t)((g.nj + LK_WARPS - 1) / LK_WARPS) * g.nk;
    if ((int64_t)grid > nitems) grid = (int)nitems;
    k_look_scan<<<grid, LK_THREADS, smem, st>>>(P);
    return 1;
}


// Patches for an output produced before the lookahead window (see k_look_patches); phi_early = that output, i fastest.
// patch_idx / patch_val / head are device buffers; returns the number of launches (0: this grid has no window).
int launch_look_patches(const uint64_t *cells, const float *phi_early, const Grid &g, bool kfastest, void *scratch, const Tuning &tun,
                        uint32_t *patch_idx, float *patch_val, uint32_t cap, unsigned int *head, cudaStream_t st)
{
    if (g.k_lo != 0 || g.k_hi != g.nk || g.slab_voxels() >= ((int64_t)1 << 32)) return 0;
    const RelaxLayout L = relax_layout(g);
    char *sc = static_cast<char *>(scratch);
    uint32_t cap_c = L.cap_c;
    if (tun.look_cap > 0) cap_c = min(cap_c, (uint32_t)tun.look_cap);
    k_look_patches<<<148, 256, 0, st>>>(cells, phi_early, g, kfastest ? 1 : 0, reinterpret_cast<const LookState *>(sc + L.look),
                                        reinterpret_cast<const uint32_t *>(sc + L.c_list), cap_c, patch_idx, patch_val, cap, head);
    return 1;
}

}  // namespace sdfb
