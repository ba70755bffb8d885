// sdfb_sweep_strips.cu -- sweep schedule "strips": warp pipelines without CTA-wide barriers.
//
// Same partial order as the other schedules (every voxel after its seven upstream neighbours,
// cpu_lib/makelevelset3.cpp:143-149), so the result is bit-identical to the reference's serial sweep; what
// changes is the unit that advances in lock step.  In sweep-relative coordinates (ri,rj,rk >= 0 counted
// from the corner the sweep starts at):
//
//   * a WARP owns 31 consecutive rows rj0 .. rj0+30 of ONE plane rk and marches along i with its lanes
//     skewed by one voxel: lane l (1..31) handles voxel ri = s - l of row rj0-1+l at step s.  Lane 0 is the
//     halo lane: it only loads the cell of row rj0-1 (left strip, or the read-only rj = 0 face).
//         m  neighbour            comes from
//         0  (ri-1, rj,   rk  )   own result of step s-1                       (register)
//         1  (ri,   rj-1, rk  )   lane l-1's result of step s-1                (shuffle)
//         2  (ri-1, rj-1, rk  )   last step's m=1                              (register)
//         3  (ri,   rj,   rk-1)   warp of plane rk-1, lane l, step s           (shared-memory ring)
//         4  (ri-1, rj,   rk-1)   last step's m=3                              (register)
//         5  (ri,   rj-1, rk-1)   lane l-1's m=3 of step s-1                   (shuffle)
//         6  (ri-1, rj-1, rk-1)   last step's m=5                              (register)
//   * a CTA is a pipeline of NW such warps on NW consecutive planes of the same strip.  Warp w publishes the
//     32 result words of each step into its ring slot and bumps a counter in shared memory; warp w+1 waits
//     on that counter (and warp w on warp w+1's, so the ring is never overrun).  There is NO CTA-wide
//     barrier in the step loop: a warp with many distance evaluations delays only its own dependents, and
//     the ring absorbs the jitter.  A feeder warp plays "plane rk0-1" for the first compute warp (words
//     loaded from global memory), a sync warp moves progress between CTAs: it publishes every warp's step
//     count to global flags (fence + store, off the compute warps' critical path) and polls the flags of the
//     strip to the left (same planes, needed by the halo lanes) and of the plane block below (feeder).
//   * strip tasks (J, KB) are handed out by an atomic ticket in anti-diagonal order J+KB to persistent
//     CTAs, so producers always hold lower tickets and are running or finished: waiting cannot deadlock.
//
// Exact pruning (stamp memo, duplicate / own-triangle rejection) and the warp-balanced evaluation queue
// are the same as in the column schedule (sdfb_sweep_columns.cu); the step itself is ~40 instructions when
// no neighbour is fresh.
#include <cstdio>
#include <cstdlib>
#include "sdfb_kernels.cuh"
#include "sdfb_sweep_common.cuh"

namespace sdfb {

namespace {

#ifndef SDFB_STRIP_NW
#define SDFB_STRIP_NW 14
#endif
#ifndef SDFB_STRIP_MINB
#define SDFB_STRIP_MINB 2
#endif
#ifndef SDFB_STRIP_SPIN
#define SDFB_STRIP_SPIN 64
#endif
#ifndef SDFB_STRIP_SLEEP
#define SDFB_STRIP_SLEEP 20
#endif
constexpr int NW = SDFB_STRIP_NW;            // compute warps (planes) per CTA
constexpr int SJ = 31;                       // rows per strip (lane 0 is the halo lane)
constexpr int RING = 8;                      // steps a producer warp may run ahead of its consumer
constexpr int FLA = 4;                       // feeder look-ahead (steps)
#ifndef SDFB_STRIP_PUBSTEP
#define SDFB_STRIP_PUBSTEP 4
#endif
constexpr int PUBSTEP = SDFB_STRIP_PUBSTEP;  // steps between progress publications to other CTAs
constexpr int GSTEP = 4;                     // steps between cta-scope fences of a compute warp (power of two, divides RING)
constexpr int NTHREADS = (NW + 2) * 32;      // feeder warp, NW compute warps, sync warp
constexpr int QCAP = 7 * 32;                 // evaluation queue entries per warp
constexpr int BIG = 0x3fffffff;
constexpr uint64_t OUT_CELL = (uint64_t)TRI_NONE;      // "no voxel here": no triangle, stamp 0

struct StripParams {
    Grid g;
    SweepDir sd;
    int rk_first, rk_last;                   // relative k range updated by this launch (inclusive)
    int NJ, NKB;                             // strips in j, plane blocks in k
    int steps;                               // steps per task: ni + 31, rounded up to a multiple of RING
    uint32_t stamp;                          // sweep_index + 1 (saturating at 31)
    uint32_t epoch;                          // progress values are epoch<<16 | steps_done
    uint8_t last[8][8];                      // memo table, see memo_last_table()
    int trace_task;
    unsigned long long *trace;               // debug (SDFB_STRIP_TRACE): per task {picked, warp0 step 0, warp0 done, last warp done} in ns
};

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// fine trace: warps 0 and 7 of one task, 8 clock64 stamps per step (region 2 of the trace buffer)
#ifdef SDFB_STRIP_FINE
#define FTRACE(slot) do { if (P.trace && lane == 0 && tk == P.trace_task && (w == 0 || w == 7)) P.trace[(size_t)65536 * 8 + ((size_t)(w ? 1 : 0) * 2048 + s) * 8 + (slot)] = clock64(); } while (0)
#else
#define FTRACE(slot) do { } while (0)
#endif

struct StripShared {
    uint32_t ring[NW + 1][RING][32];         // [0] feeder, [w+1] compute warp w
    uint32_t q_ent[NW][QCAP];                // (owner lane << 27) | tri
    float q_d[NW][QCAP];
    uint32_t thr[8][8];                      // memo thresholds per voxel class and offset
    unsigned long long full[NW + 1][RING];   // mbarriers: ring slot written (1 arrival: the producer's lane 0)
    unsigned long long empty[NW + 1][RING];  // mbarriers: ring slot read    (1 arrival: the consumer's lane 0)
    int cnt[NW + 1];                         // steps completed: [0] feeder, [w+1] compute warp w (fast-path hint)
    int gcnt[NW];                            // steps whose cell stores are fenced to cta scope (read by the sync warp)
    int left_ok[NW];                         // steps the left strip's warp w is known to have completed
    int down_ok;                             // steps the last warp of the plane block below has completed
    int task;
};

// Waiting for progress published by the sync warp (other CTAs' flags): these waits are long while a task's
// producers are still far behind, so back off instead of burning issue slots.
__device__ __forceinline__ void wait_ge(int &seen, const int *p, int need)
{
    if (seen >= need) return;
    unsigned ns = 32;
    for (int it = 0;; ++it) {
        seen = *reinterpret_cast<const volatile int *>(p);
        if (seen >= need) break;
        if (it >= SDFB_STRIP_SPIN) { __nanosleep(ns); if (ns < 1024) ns <<= 1; }
    }
    asm volatile("" ::: "memory");
}

// ---- mbarrier helpers (shared::cta, one arrival per phase) -----------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Blocks (hardware sleep, bounded) until the phase with the given parity has completed.
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
// Ring hand-over between adjacent warps of the pipeline.  `seen` caches the partner's step counter (a plain
// word in shared memory): the common case is a register compare.  When the partner is not there yet the warp
// sleeps on the slot's mbarrier instead of polling.  g = running step index of this CTA (never reset, so the
// mbarrier phases continue across tasks), q = g / RING = phase index of slot g % RING.
__device__ __forceinline__ void ring_wait(int &seen, const int *cnt, int need, unsigned long long *bar, uint32_t parity)
{
    if (seen >= need) return;
    seen = *reinterpret_cast<const volatile int *>(cnt);
    if (seen >= need) { asm volatile("" ::: "memory"); return; }
    mbar_wait(bar, parity);
    seen = need;
}

// L2 prefetch of a run of 64 cells of one row with one bulk request (no data returned); see the column schedule.
constexpr int PF_CELLS = 64, PF_AHEAD = 96;
__device__ __forceinline__ void prefetch_run(const uint64_t *row0, int64_t si, int first, int ni)
{
    int lo = first, hi = first + PF_CELLS - 1;
    if (hi > ni - 1) hi = ni - 1;
    if (lo < 0) lo = 0;
    if (lo > hi) return;
    const uint64_t *a = row0 + si * (int64_t)lo, *b = row0 + si * (int64_t)hi;
    const uint64_t *beg = si > 0 ? a : b;
    uintptr_t addr = reinterpret_cast<uintptr_t>(beg) & ~(uintptr_t)15;
    uint32_t bytes = (uint32_t)((hi - lo + 1) * 8 + 16) & ~15u;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(addr), "r"(bytes) : "memory");
}

// ---- feeder warp: plane rk0-1 (the plane block below, a read-only face or a slab halo plane) -------------
__device__ __forceinline__ void feeder_strip(const uint64_t *__restrict__ cells, const StripParams &P, StripShared &sh,
                                             int lane, int J, int KB, int rj0, int rk0, int gbase)
{
    const Grid &g = P.g;
    const int rj = rj0 - 1 + lane, rk = rk0 - 1;
    const bool row_ok = rj <= g.nj - 1;
    const int64_t si = (int64_t)P.sd.di;
    const uint64_t *row0 = cells;
    if (row_ok) row0 = cells + g.cidx(P.sd.abs_i(0, g), P.sd.abs_j(rj, g), P.sd.abs_k(rk, g));
    int down_seen = (KB > 0) ? 0 : BIG, left_seen = (KB > 0 && J > 0) ? 0 : BIG, cons_seen = 0;
    uint32_t wd[FLA];
    wait_ge(down_seen, &sh.down_ok, min(P.steps, FLA));
    wait_ge(left_seen, &sh.left_ok[0], min(P.steps, FLA - 1 + 32));
    if (row_ok) { prefetch_run(row0, si, 0, g.ni); prefetch_run(row0, si, PF_CELLS, g.ni); }
    #pragma unroll
    for (int u = 0; u < FLA; ++u) {
        const int ri = u - lane;
        wd[u] = TRI_NONE;
        if (row_ok && (unsigned)ri < (unsigned)g.ni) wd[u] = __ldcg(reinterpret_cast<const uint32_t *>(row0 + si * (int64_t)ri));
    }
    for (int s0 = 0; s0 < P.steps; s0 += FLA) {
        #pragma unroll
        for (int u = 0; u < FLA; ++u) {
            const int s = s0 + u;
            const int gs = gbase + s, slot = gs & (RING - 1), q = gs / RING;
            if (s >= RING) ring_wait(cons_seen, &sh.cnt[1], s - RING + 1, &sh.empty[0][slot], (uint32_t)(q - 1) & 1u);
            sh.ring[0][slot][lane] = wd[u];
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&sh.full[0][slot]);
                *reinterpret_cast<volatile int *>(&sh.cnt[0]) = s + 1;
            }
            const int t = s + FLA;
            wd[u] = TRI_NONE;
            if (t < P.steps) {
                // the word of step t was produced by the block below at its step t (row rj0-1: by the diagonal
                // task at step t+31, which the left strip's first warp has consumed by its own step t+31)
                wait_ge(down_seen, &sh.down_ok, min(P.steps, t + 1));
                wait_ge(left_seen, &sh.left_ok[0], min(P.steps, t + 32));
                const int ri = t - lane;
                if (row_ok && (unsigned)ri < (unsigned)g.ni) wd[u] = __ldcg(reinterpret_cast<const uint32_t *>(row0 + si * (int64_t)ri));
                if (row_ok && (t & (PF_CELLS - 1)) == 0) prefetch_run(row0, si, ri + PF_AHEAD, g.ni);
            }
        }
    }
}

// ---- sync warp: lane w < NW serves compute warp w, lane NW polls the block below ---------------------------
__device__ __forceinline__ void sync_strip(const StripParams &P, StripShared &sh, int lane, int J, int KB,
                                           uint32_t *__restrict__ progress)
{
    const uint32_t ebase = P.epoch << 16;
    const bool active = lane < NW;
    const bool has_left = active && J > 0, has_down = (lane == NW) && KB > 0;
    uint32_t *mine = progress + ((size_t)(KB * P.NJ + J) * NW + (active ? lane : 0));
    const uint32_t *left = progress + ((size_t)(KB * P.NJ + (J > 0 ? J - 1 : 0)) * NW + (active ? lane : 0));
    const uint32_t *down = progress + ((size_t)((KB > 0 ? KB - 1 : 0) * P.NJ + J) * NW + (NW - 1));
    int pub = 0;
    for (;;) {
        int c = P.steps;
        if (active) c = *reinterpret_cast<const volatile int *>(&sh.gcnt[lane]);
        // Publish in chunks of PUBSTEP steps: a gpu-scope fence also invalidates the SM's L1 (CCTL.IVALL), which
        // the triangle-record gathers live on, so it must stay rare.
        const bool want = active && (c >= pub + PUBSTEP || (c >= P.steps && pub < P.steps));
        if (__any_sync(0xffffffffu, want)) {
            __threadfence();                 // release: the cells stored before the counters were read ...
            if (active && c > pub) { *reinterpret_cast<volatile uint32_t *>(mine) = ebase + (uint32_t)c; pub = c; }   // ... precede the flag
        }
        if (has_left) {
            const uint32_t v = *reinterpret_cast<const volatile uint32_t *>(left);
            *reinterpret_cast<volatile int *>(&sh.left_ok[lane]) = v >= ebase ? (int)(v - ebase) : 0;
        }
        if (has_down) {
            const uint32_t v = *reinterpret_cast<const volatile uint32_t *>(down);
            *reinterpret_cast<volatile int *>(&sh.down_ok) = v >= ebase ? (int)(v - ebase) : 0;
        }
        if (__all_sync(0xffffffffu, !active || pub >= P.steps)) break;
    }
}

// ---- compute warps -------------------------------------------------------------------------------------------
struct StripLane {
    uint64_t *ptr;                            // cell of this step's voxel (row0 + si*ri)
    int ri;
    uint32_t prev_lo, m1_old, m3_old, m5_old; // rolled neighbour words (see the table at the top)
    int prod_seen, cons_seen, left_seen;
    unsigned changed, evals;
};

// The rare part of a step: at least one lane has a neighbour that changed since the lane last looked at that
// offset.  Filter, queue, evaluate round-robin over the warp, replay in the reference's order with strict "<".
__device__ __forceinline__ uint2 strip_evaluate(const TriRec *__restrict__ rec, const StripParams &P, StripShared &sh,
                                                int w, int lane, int s, int rj0, float gz, bool update, int cls,
                                                uint32_t nb0, uint32_t nb1, uint32_t nb2, uint32_t nb3, uint32_t nb4,
                                                uint32_t nb5, uint32_t nb6, uint32_t cur, uint64_t *self_ptr, float phi)
{
    const Grid &g = P.g;
    uint32_t *const q_ent = sh.q_ent[w];
    float *const q_d = sh.q_d[w];
    const uint32_t nb[7] = {nb0, nb1, nb2, nb3, nb4, nb5, nb6};
    uint32_t live = 0;
    if (update) {
        #pragma unroll
        for (int m = 0; m < 7; ++m) {
            const uint32_t x = nb[m];
            const bool keep = ((x & TRI_MASK) != TRI_NONE) && (((x ^ cur) & TRI_MASK) != 0) && (x >= sh.thr[cls][m]);
            live |= keep ? (1u << m) : 0u;
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
    if ((b0 | b1 | b2) == 0) return make_uint2(cur, 0u);
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int total = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
    const int off = __popc(b0 & lt_mask) + 2 * __popc(b1 & lt_mask) + 4 * __popc(b2 & lt_mask);
    if (live) {
        int q = off;
        #pragma unroll
        for (int m = 0; m < 7; ++m) if ((live >> m) & 1u) {
            q_ent[q] = ((uint32_t)lane << 27) | (nb[m] & TRI_MASK); ++q;
            const char *ra = reinterpret_cast<const char *>(&rec[nb[m] & TRI_MASK]);     // start the gather now
            asm volatile("prefetch.global.L1 [%0];" ::"l"(ra));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(ra + 32));
        }
    }
    __syncwarp();
    unsigned evals = 0, changed = 0;
    for (int q = lane; q < total; q += 32) {
        const uint32_t e = q_ent[q];
        const int ol = (int)(e >> 27);                                // owner lane -> its voxel's position
        const F3 gx{lattice(P.sd.abs_i(s - ol, g), g.dx, g.ox), lattice(P.sd.abs_j(rj0 - 1 + ol, g), g.dx, g.oy), gz};
        const TriRec *tr = &rec[e & TRI_MASK];
        const float4 p = __ldg(&tr->p), qq = __ldg(&tr->q), r = __ldg(&tr->r);
        q_d[q] = ptd_rec(gx, p, qq, r);
        ++evals;
    }
    __syncwarp();
    if (live) {
        uint32_t best = TRI_NONE;
        for (int q = off; q < off + ncand; ++q) {                     // the reference's order and strict "<"
            const float d = q_d[q];
            if (d < phi) { phi = d; best = q_ent[q] & TRI_MASK; }
        }
        if (best != TRI_NONE) {
            cur = (P.stamp << 27) | best;
            *self_ptr = pack_cell(phi, cur);
            changed = 1;
        }
    }
    __syncwarp();
    return make_uint2(cur, (evals << 1) | changed);
}

// One step of a compute warp.  PAR = step parity; `own` holds the lane's cell of this step on entry and is
// reloaded with the cell two steps ahead (never copied: a copy of an in-flight load would stall on it).
template <int PAR>
__device__ __forceinline__ void strip_step(const TriRec *__restrict__ rec, const StripParams &P, StripShared &sh,
                                           int w, int lane, int s, int gbase, int tk, int rj0, float gz, bool row_ok, int row_class,
                                           uint32_t tmin, uint32_t tmin_e, bool has_cons, const uint64_t *row0,
                                           uint64_t &own, StripLane &st)
{
    const int ni = P.g.ni;
    const int64_t si = (int64_t)P.sd.di;
    const int ri = st.ri;
    // the plane below must have published step s; our consumer must have drained the ring slot we overwrite
    const int gs = gbase + s, slot = gs & (RING - 1), q = gs / RING;
    FTRACE(0);
    ring_wait(st.prod_seen, &sh.cnt[w], s + 1, &sh.full[w][slot], (uint32_t)q & 1u);
    FTRACE(1);
    if (has_cons && s >= RING) ring_wait(st.cons_seen, &sh.cnt[w + 2], s - RING + 1, &sh.empty[w + 1][slot], (uint32_t)(q - 1) & 1u);
    FTRACE(2);
    const uint32_t m3 = sh.ring[w][slot][lane];
    const uint32_t m1 = __shfl_up_sync(0xffffffffu, st.prev_lo, 1);
    const uint32_t m5 = __shfl_up_sync(0xffffffffu, st.m3_old, 1);
    const uint64_t self = own;
    uint64_t *const self_ptr = st.ptr;
    // reload: the cell of step s+2.  The halo lane's cell is produced by the left strip's lane 31 at its step
    // s+2+31 and is read from L2; own rows are only ever written by this lane (L1 is fine).
    FTRACE(3);
    wait_ge(st.left_seen, &sh.left_ok[w], min(P.steps, s + 2 + 32));
    FTRACE(4);
    own = OUT_CELL;
    if (row_ok && (unsigned)(ri + 2) < (unsigned)ni) {
        if (lane == 0) own = __ldcg(self_ptr + 2 * si);
        else own = *(self_ptr + 2 * si);
    }
    if (row_ok && (s & (PF_CELLS - 1)) == 0) prefetch_run(row0, si, ri + PF_AHEAD, ni);
    uint32_t cur = cell_lo(self);
    const bool update = row_ok && lane > 0 && (unsigned)(ri - 1) < (unsigned)(ni - 1);     // 1 <= ri <= ni-1
    const bool edge = (ri == ni - 1);
    // cheap test first: no neighbour word reaches the lowest memo threshold -> nothing to evaluate
    const uint32_t mx = max(max(max(st.prev_lo, m1), max(st.m1_old, m3)), max(max(st.m3_old, m5), st.m5_old));
    const bool fresh = update && mx >= (edge ? tmin_e : tmin);
#ifdef SDFB_STRIP_FINE
    if (P.trace && tk == P.trace_task && (w == 0 || w == 7)) { if (cur == 0x12345678u) st.changed += 1; FTRACE(5); }   // forces the own load to have landed
#endif
    if (__any_sync(0xffffffffu, fresh)) {
        const uint2 r = strip_evaluate(rec, P, sh, w, lane, s, rj0, gz, fresh, row_class | (edge ? 1 : 0),
                                       st.prev_lo, m1, st.m1_old, m3, st.m3_old, m5, st.m5_old,
                                       cur, self_ptr, cell_phi(self));
        cur = r.x; st.changed += r.y & 1u; st.evals += r.y >> 1;
    }
    sh.ring[w + 1][slot][lane] = cur;
    st.m1_old = m1; st.m5_old = m5; st.m3_old = m3; st.prev_lo = cur;
    st.ptr = self_ptr + si;
    st.ri = ri + 1;
    FTRACE(6);
    __syncwarp();
    if (lane == 0) {
        mbar_arrive(&sh.empty[w][slot]);     // the ring slot of the plane below has been read by all lanes
        if (has_cons) mbar_arrive(&sh.full[w + 1][slot]);
        *reinterpret_cast<volatile int *>(&sh.cnt[w + 1]) = s + 1;
        if ((s & (GSTEP - 1)) == GSTEP - 1) {
            __threadfence_block();           // the warp's cell stores precede the counter (the sync warp fences to gpu scope)
            *reinterpret_cast<volatile int *>(&sh.gcnt[w]) = s + 1;
        }
    }
    FTRACE(7);
}

__device__ __forceinline__ void compute_strip(uint64_t *__restrict__ cells, const TriRec *__restrict__ rec,
                                              const StripParams &P, StripShared &sh, int w, int lane, int J, int rj0,
                                              int rk, int gbase, int tk, unsigned &my_changed, unsigned &my_evals)
{
    const Grid &g = P.g;
    const int rj = rj0 - 1 + lane;
    // a warp whose plane lies beyond the last one still takes part in the pipeline (it passes "no voxel" on)
    const bool row_ok = rj <= g.nj - 1 && rk <= P.rk_last;
    const int64_t si = (int64_t)P.sd.di;
    const int k = P.sd.abs_k(min(rk, P.rk_last), g);
    const float gz = lattice(k, g.dx, g.oz);
    uint64_t *row0 = cells;
    if (row_ok) row0 = cells + g.cidx(P.sd.abs_i(0, g), P.sd.abs_j(rj, g), k);
    const int row_class = (rj == g.nj - 1 ? 2 : 0) | (rk == g.nk - 1 ? 4 : 0);
    uint32_t tmin = 0xffffffffu, tmin_e = 0xffffffffu;
    #pragma unroll
    for (int m = 0; m < 7; ++m) { tmin = min(tmin, sh.thr[row_class][m]); tmin_e = min(tmin_e, sh.thr[row_class | 1][m]); }
    StripLane st;
    st.ri = -lane;
    st.ptr = row0 + si * (int64_t)st.ri;
    st.prev_lo = TRI_NONE; st.m1_old = TRI_NONE; st.m3_old = TRI_NONE; st.m5_old = TRI_NONE;
    st.prod_seen = 0; st.cons_seen = 0; st.left_seen = (J > 0 && rk <= P.rk_last) ? 0 : BIG;
    st.changed = 0; st.evals = 0;
    const bool has_cons = (w + 1 < NW);
    uint64_t ownA = OUT_CELL, ownB = OUT_CELL;
    wait_ge(st.left_seen, &sh.left_ok[w], min(P.steps, 1 + 32));
    if (row_ok) {
        prefetch_run(row0, si, 0, g.ni);
        prefetch_run(row0, si, PF_CELLS, g.ni);
        if ((unsigned)st.ri < (unsigned)g.ni) ownA = (lane == 0) ? __ldcg(st.ptr) : *st.ptr;
        if ((unsigned)(st.ri + 1) < (unsigned)g.ni) ownB = (lane == 0) ? __ldcg(st.ptr + si) : *(st.ptr + si);
    }
    for (int s = 0; s < P.steps; s += 2) {
        if (P.trace && lane == 0 && w == 0 && (s == 2 || s == 66)) P.trace[(size_t)tk * 8 + (s == 2 ? 1 : 4)] = gtime();
        strip_step<0>(rec, P, sh, w, lane, s, gbase, tk, rj0, gz, row_ok, row_class, tmin, tmin_e, has_cons, row0, ownA, st);
        strip_step<1>(rec, P, sh, w, lane, s + 1, gbase, tk, rj0, gz, row_ok, row_class, tmin, tmin_e, has_cons, row0, ownB, st);
    }
    if (P.trace && lane == 0 && (w == 0 || w == NW - 1)) P.trace[(size_t)tk * 8 + (w == 0 ? 2 : 3)] = gtime();
    my_changed += st.changed; my_evals += st.evals;
}

__global__ void __launch_bounds__(NTHREADS, SDFB_STRIP_MINB)
k_sweep_strips(uint64_t *__restrict__ cells, const TriRec *__restrict__ rec, StripParams P,
               uint32_t *__restrict__ progress, uint32_t *__restrict__ ticket, unsigned long long *__restrict__ changed)
{
    __shared__ StripShared sh;
    const int tid = threadIdx.x, lane = tid & 31, wi = tid >> 5;
    const int ntasks = P.NJ * P.NKB;
    unsigned my_changed = 0, my_evals = 0;
    if (tid < 64) {
        const uint32_t l = P.last[tid >> 3][tid & 7];
        sh.thr[tid >> 3][tid & 7] = ((tid & 7) == 7) ? 0xffffffffu : (l ? (l + 1u) << 27 : 0u);
    }
    if (tid < (NW + 1) * RING) { mbar_init(&sh.full[0][0] + tid, 1); mbar_init(&sh.empty[0][0] + tid, 1); }
    int gbase = 0;                            // running step index: the mbarrier phases continue across tasks
    for (;; gbase += P.steps) {
        if (tid == 0) sh.task = (int)atomicAdd(ticket, 1u);
        __syncthreads();
        const int tk = sh.task;
        if (tk >= ntasks) break;
        int J, KB;
        {
            int d = 0, rem = tk;
            for (;;) {
                const int lo = max(0, d - (P.NKB - 1)), hi = min(d, P.NJ - 1);
                const int cnt = hi - lo + 1;
                if (rem < cnt) { J = lo + rem; KB = d - J; break; }
                rem -= cnt; ++d;
            }
        }
        const int rj0 = 1 + J * SJ, rk0 = P.rk_first + KB * NW;
        if (tid <= NW) sh.cnt[tid] = 0;
        if (tid < NW) sh.gcnt[tid] = 0;
        if (P.trace && tid == 0) { P.trace[(size_t)tk * 8] = gtime(); P.trace[(size_t)tk * 8 + 5] = (unsigned long long)J << 32 | (unsigned)KB; P.trace[(size_t)tk * 8 + 6] = blockIdx.x; }
        if (tid < NW) sh.left_ok[tid] = (J > 0) ? 0 : BIG;
        if (tid == 0) sh.down_ok = (KB > 0) ? 0 : BIG;
        __syncthreads();
        if (wi == 0) {
            feeder_strip(cells, P, sh, lane, J, KB, rj0, rk0, gbase);
        } else if (wi <= NW) {
            compute_strip(cells, rec, P, sh, wi - 1, lane, J, rj0, rk0 + wi - 1, gbase, tk, my_changed, my_evals);
        } else {
            sync_strip(P, sh, lane, J, KB, progress);
        }
        __syncthreads();
    }
    // ---- teardown: count changes; the last CTA out resets the ticket for the next launch ----------
    unsigned wsum = my_changed, esum = my_evals;
    for (int o = 16; o > 0; o >>= 1) { wsum += __shfl_down_sync(0xffffffffu, wsum, o); esum += __shfl_down_sync(0xffffffffu, esum, o); }
    if (lane == 0 && wsum) atomicAdd(changed, (unsigned long long)wsum);
    if (lane == 0 && esum) atomicAdd(changed + 1, (unsigned long long)esum);
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned done = atomicAdd(ticket + 1, 1u);
        if (done == gridDim.x - 1) { ticket[0] = 0; ticket[1] = 0; __threadfence(); }
    }
}

}  // namespace

// progress: 4 words (ticket, exit count, spare) + one flag per (task, warp)
size_t sweep_strips_progress_words(const Grid &g)
{
    const size_t NJ = (size_t)(g.nj + SJ - 1) / SJ + 1, NKB = (size_t)(g.nkl() + NW - 1) / NW + 1;
    return 4 + NJ * NKB * NW;
}

int launch_sweep_strips(uint64_t *cells, const TriRec *rec, const Grid &g, int sweep_index,
                        unsigned long long *changed, uint32_t *progress, uint32_t epoch, cudaStream_t st)
{
    StripParams P{};
    P.g = g;
    P.sd = SweepDir::of(sweep_index);
    int rk_lo, rk_hi;
    if (!P.sd.owned_rk_range(g, rk_lo, rk_hi)) return 0;
    if (g.ni < 2 || g.nj < 2) return 0;
    P.rk_first = rk_lo; P.rk_last = rk_hi;
    P.NJ = (g.nj - 1 + SJ - 1) / SJ;
    P.NKB = (rk_hi - rk_lo + 1 + NW - 1) / NW;
    P.steps = (g.ni + 31 + RING - 1) / RING * RING;    // whole ring turns: slot and phase of a step follow from the running index
    P.stamp = (uint32_t)min(sweep_index + 1, 31);
    P.epoch = epoch;
    memo_last_table(sweep_index, P.sd, P.last);
    int dev = 0, sms = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sweep_strips, NTHREADS, 0);
    if (occ < 1) occ = 1;
    if (getenv("SDFB_MAX_OCC")) occ = min(occ, atoi(getenv("SDFB_MAX_OCC")));   // experiment knob
    int grid = sms * occ;
    const int ntasks = P.NJ * P.NKB;
    if (grid > ntasks) grid = ntasks;
    static unsigned long long *trace_buf = nullptr;
    const char *trace_path = getenv("SDFB_STRIP_TRACE");
    if (trace_path) {
        if (!trace_buf) cudaMalloc(&trace_buf, (size_t)(65536 + 4096) * 8 * sizeof(unsigned long long));
        cudaMemsetAsync(trace_buf, 0, (size_t)(65536 + 4096) * 8 * sizeof(unsigned long long), st);
        if (ntasks <= 65536 && P.steps <= 2048) P.trace = trace_buf;
        P.trace_task = getenv("SDFB_STRIP_TRACE_TASK") ? atoi(getenv("SDFB_STRIP_TRACE_TASK")) : ntasks / 2;
    }
    k_sweep_strips<<<grid, NTHREADS, 0, st>>>(cells, rec, P, progress + 4, progress, changed);
    if (P.trace) {
        cudaStreamSynchronize(st);
        unsigned long long *h = (unsigned long long *)malloc((size_t)ntasks * 8 * sizeof(unsigned long long));
        cudaMemcpy(h, trace_buf, (size_t)ntasks * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        char name[512];
        snprintf(name, sizeof(name), "%s.%d.bin", trace_path, sweep_index);
        FILE *f = fopen(name, "wb");
        if (f) { fwrite(h, sizeof(unsigned long long), (size_t)ntasks * 8, f); fclose(f); }
        free(h);
#ifdef SDFB_STRIP_FINE
        h = (unsigned long long *)malloc((size_t)4096 * 8 * sizeof(unsigned long long));
        cudaMemcpy(h, trace_buf + (size_t)65536 * 8, (size_t)4096 * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        snprintf(name, sizeof(name), "%s.%d.fine.bin", trace_path, sweep_index);
        f = fopen(name, "wb");
        if (f) { fwrite(h, sizeof(unsigned long long), (size_t)4096 * 8, f); fclose(f); }
        free(h);
#endif
    }
    return 1;
}

}  // namespace sdfb
