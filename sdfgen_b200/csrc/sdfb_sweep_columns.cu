// sdfb_sweep_columns.cu -- the production sweep schedule: pipelined columns.
//
// One launch per sweep direction (k_sweep_columns), or several sweeps in one launch with consecutive sweeps overlapping
// where their directions allow it (k_sweep_columns_fused, near the end of this file): the first pass of a plan on one GPU,
// all 16 sweeps of a LINKED k-slab -- the exact multi-GPU mode, in which the columns of a slab's last K block hand their
// boundary-plane cells to the slab above over NVLink (template parameter LINK).
// In sweep-relative coordinates (ri,rj,rk >= 0, counted from the corner the sweep starts at) the (rj,rk) plane is cut into columns of EJ x EK rows.  A CTA owns one
// column at a time and marches along i: lane (a,b) of the column handles voxel ri = s - a - b - 2 at step
// s, so lanes are skewed along the anti-diagonal and every one of the seven upstream neighbours
// (cpu_lib/makelevelset3.cpp:143-149) was produced 1..3 steps earlier by lane (a-1|a, b-1|b):
//
//      m  neighbour (rel.)        produced by lane   at step
//      0  (ri-1, rj,   rk  )      (a,   b  )         s-1   (register)
//      1  (ri,   rj-1, rk  )      (a-1, b  )         s-1
//      2  (ri-1, rj-1, rk  )      (a-1, b  )         s-2
//      3  (ri,   rj,   rk-1)      (a,   b-1)         s-1
//      4  (ri-1, rj,   rk-1)      (a,   b-1)         s-2
//      5  (ri,   rj-1, rk-1)      (a-1, b-1)         s-2
//      6  (ri-1, rj-1, rk-1)      (a-1, b-1)         s-3
//
// The lanes exchange the 32-bit {stamp|closest_tri} words through a double-buffered array in shared
// memory (three reads per step, all of step s-1; older words roll through registers), one
// __syncthreads per step.  Lanes a=-1 / b=-1 are "halo lanes": two extra warps that, instead of
// computing, load the word of their voxel from global memory -- it belongs to the column to the left
// (J-1), below (K-1) or diagonal, or to the read-only ri/rj/rk = 0 faces (or a slab halo plane).
// Columns are handed out by an atomic ticket in anti-diagonal order (J+K), so a column's producers
// always hold lower tickets and are running or finished: the spin on their progress counters cannot
// deadlock.  A column publishes its step count every PUBLISH steps (CTA barrier, then a RELEASE store by the
// sync warp: st.release.gpu, not __threadfence() + store -- fence.sc drags a CCTL.IVALL behind it that empties
// the SM's L1, the cache the triangle records live in); a consumer column may run step s once
// left >= s+EJ+2+... (see sync_column).
// Every dependency of the serial Gauss-Seidel order is respected, so the result is bit-identical to
// the reference's single-threaded sweep.
//
// Work per voxel is data dependent (0..7 distance evaluations), so evaluation is decoupled from
// ownership: each lane filters its candidates (drop "no triangle", the voxel's own triangle,
// duplicates, and -- the stamp memo -- neighbours whose triangle has not changed since the last sweep
// in which this voxel looked at the same offset: a candidate that lost once can never win later
// because phi only decreases), pushes the survivors to a per-warp queue in shared memory, the 32
// lanes evaluate the queue round-robin, and each owner then replays its own results in the
// reference's order with the reference's strict "<".
#include <cstdio>
#include <cstdlib>
#include "sdfb_kernels.cuh"
#include "sdfb_sweep_common.cuh"

// This file is compiled TWICE into the library: as it stands (8 x 16 columns: CTAs of 192 threads, three or four per SM), and
// through sdfb_sweep_columns_ek12.cu with 8 x 12 columns (CTAs of 160 threads, FOUR per SM at the full 93 registers), whose
// entry points carry the suffix _ek12.  Small launches are bound by the length of the wavefront's dependency chain and by
// how many columns are resident, and run 5-6 % faster with the smaller CTAs (C2 first pass 54.6 -> 51.5 ms); large ones
// (>= 300 M voxels) are throughput-bound and keep 8 x 16 (C3: 232 ms against 256).  sdfb_api.cu chooses by launch size.
#ifndef SDFB_COLS_SUFFIX
#define SDFB_COLS_SUFFIX
#endif
#define SDFB_COLS_CAT2(a, b) a##b
#define SDFB_COLS_CAT(a, b) SDFB_COLS_CAT2(a, b)
#define COLS_FN(name) SDFB_COLS_CAT(name, SDFB_COLS_SUFFIX)

namespace sdfb {

namespace {

// 8 x 16 columns at 3-4 CTAs per SM measured ~6 % faster than 16 x 16 at 2 (more independent barrier domains);
// 16 x 8 and 8 x 8 were slower.  Two queue entries per lane and trip (ILP) measured 25 % SLOWER, with or without spills.
#ifndef SDFB_EJ
#define SDFB_EJ 8
#endif
#ifndef SDFB_EK
#define SDFB_EK 16
#endif
constexpr int EJ = SDFB_EJ, EK = SDFB_EK;    // column extent (rows x planes)
constexpr int NCOMPUTE = EJ * EK;            // compute lanes (16 x 16: 8 warps)
constexpr int NHALO = (EJ + EK + 1 + 31) / 32 * 32;   // halo lanes, whole warps
constexpr int NSTEPPERS = NCOMPUTE + NHALO;  // lanes that take part in the per-step barrier
// SDFB_WG: warpgroup register split.  The CTA is padded to two warpgroups (compute warps | halo, sync and two idle
// warps) and compiled for 64 registers per thread, so that four CTAs fit an SM; the compute warpgroup then takes
// the registers the other one gives up (setmaxnreg is a warpgroup-wide operation).  The halo warp does not
// evaluate in this layout.
#ifndef SDFB_WG
#define SDFB_WG 0
#endif
#ifndef SDFB_WG_COMPUTE_REGS
#define SDFB_WG_COMPUTE_REGS 96
#endif
#ifndef SDFB_WG_LIGHT_REGS
#define SDFB_WG_LIGHT_REGS 32
#endif
#ifndef SDFB_WG_HALO_EVAL
#define SDFB_WG_HALO_EVAL 0
#endif
constexpr bool WG = SDFB_WG != 0;
constexpr bool HALO_EVAL = !WG || SDFB_WG_HALO_EVAL != 0;   // the halo warp takes a share of the evaluations
static_assert(!WG || NCOMPUTE == 128, "the warpgroup layout needs exactly four compute warps");
static_assert(EJ * EK <= 256, "queue entries carry the owner lane in 8 bits");
// SDFB_MERGE_SYNC: no separate sync warp -- the halo warp polls its producers' progress words itself (only when the values
// it last saw do not cover the next chunk) and publishes the column's progress right after the chunk's last step barrier.
// A CTA is 32 threads smaller, so one more fits an SM at the full register count (8 x 12: five of 128 threads, 8 x 16: four of
// 160 without spills), and the polling loops (a third of the executed instructions) disappear.
#ifndef SDFB_MERGE_SYNC
#define SDFB_MERGE_SYNC 0
#endif
constexpr bool MERGE = SDFB_MERGE_SYNC != 0 && !WG;
constexpr int NTHREADS = WG ? 256 : (MERGE ? NSTEPPERS : NSTEPPERS + 32);     // + 1 sync warp (flag polling and progress publication)
constexpr int EVAL_LANES = HALO_EVAL ? NSTEPPERS : NCOMPUTE;   // lanes that share the column-wide evaluation queue
#ifndef SDFB_SYNC_SLEEP
#define SDFB_SYNC_SLEEP 150
#endif
#ifndef SDFB_PUBLISH
#define SDFB_PUBLISH 2
#endif
constexpr int PUBLISH = SDFB_PUBLISH;                   // steps between progress publications
// The kernels are built for two register bounds, template parameter MINB = 3 and 4 CTAs per SM; SDFB_MINB_HI replaces the
// "4" (smaller column cross-sections make smaller CTAs, of which more fit an SM).
#ifndef SDFB_MINB_HI
#define SDFB_MINB_HI 4
#endif
constexpr int minb_bound(int m) { return m >= 4 ? SDFB_MINB_HI : m; }
// register bound for launches below 300 M voxels (the 160-thread CTAs of the 8 x 12 build fit four to an SM without spills)
#ifndef SDFB_MINB_SMALL
#define SDFB_MINB_SMALL 3
#endif
constexpr int MINB_SMALL = SDFB_MINB_SMALL;
// SDFB_QATOMIC: the warps of a column reserve their slice of the column-wide queue with one shared-memory atomicAdd
// instead of exchanging their totals through a barrier: a step with candidates takes 3 CTA barriers instead of 4.
// The order of the slices in the queue then depends on arrival order, which only changes WHICH lane evaluates an
// entry; every owner still replays its own entries in the reference's order.
#ifndef SDFB_QATOMIC
#define SDFB_QATOMIC 0
#endif
constexpr bool QATOMIC = SDFB_QATOMIC != 0;
// SDFB_ONEBAR: every warp of a column enqueues into ITS OWN region of the queue (offsets from its own ballots) before the
// first queue barrier; the evaluating lanes map a global entry number to (warp region, local index) from the four warp
// totals.  The barrier that only existed to exchange those totals before the enqueue disappears: a working step takes 3
// CTA barriers instead of 4 and the enqueue no longer waits for the slowest warp's filter.  No atomic (unlike
// SDFB_QATOMIC): nothing new on any warp's critical path but three compares per evaluated entry.
#ifndef SDFB_ONEBAR
#define SDFB_ONEBAR 0
#endif
constexpr bool ONEBAR = SDFB_ONEBAR != 0;
static_assert(!(ONEBAR && QATOMIC), "SDFB_ONEBAR and SDFB_QATOMIC are alternatives");
// SDFB_EVAL_PF: before a lane evaluates a queue entry it starts the record gather of its NEXT entry (L1 prefetch), so
// the second evaluation round of a step finds its triangle in L1.  SDFB_ENQ_PF: the enqueuing lane prefetches the
// records of its own candidates into the SM's L1 (every lane of the column shares it) two barriers before they are read.
#ifndef SDFB_EVAL_PF
#define SDFB_EVAL_PF 0
#endif
#ifndef SDFB_ENQ_PF
#define SDFB_ENQ_PF 0
#endif
// SDFB_STAGE (experiment of round 2, OFF): triangle records of a step's candidates are copied into shared memory with
// cp.async (LDGSTS) by the lane that owns the voxel as soon as its candidate list is known, so that the record gather
// overlaps the queue barriers instead of heading every evaluation round (profiles/r2_column_step_phases.txt: one round =
// ~1000 cycles for ~190 instructions).  Bit-exact and 16-24 % SLOWER at C2 (first pass 67-74 ms with 2-5 slots per lane
// against 54.9 ms without, profiles/r2_variants.txt): neighbouring voxels share candidates, so the L1-cached gather by
// the evaluating lanes moves far fewer bytes than one private copy per (voxel, candidate), and the copy instructions sit
// on the filter's critical path.  -DSDFB_STAGE=<slots per lane> builds it.
#ifndef SDFB_STAGE
#define SDFB_STAGE 0
#endif
constexpr int STAGE = SDFB_STAGE;            // staged records per compute lane and step (0 = off)
constexpr int SHIFT = 2;                     // lane (a,b) handles ri = s - a - b - SHIFT, so halo lane (-1,-1) starts at ri = 0
constexpr int QCAP = 7 * 32;                 // queue entries per warp

struct ColParams {
    Grid g;
    SweepDir sd;
    int rk_first, rk_last;                   // relative k range updated by this launch (inclusive)
    int NJ, NK;                              // columns in j and k
    int steps;                               // steps per column = ni + EJ + EK, rounded up to even
    uint32_t stamp;                          // sweep_index + 1
    uint32_t epoch;                          // progress values are epoch<<16 | steps_done
    int trace_col;                           // ticket of the column to trace
    const unsigned int *run_if;              // if set: the launch does nothing unless this word is non-zero (the
                                             // relaxation schedule's fallback, see sdfb_sweep_relax.cu)
    unsigned long long *trace;               // debug builds (-DSDFB_TRACE) only: per-warp step timestamps
    // last[c][m]: stamp (sweep index + 1) of the latest earlier sweep in which a voxel of class c examined
    // neighbour offset m, 0 if none.  Class bits: 1 = last voxel of its row (ri = ni-1), 2 = last row
    // (rj = nj-1), 4 = last plane (rk = nk-1): such voxels lie on a grid face and are only visited by sweeps
    // whose direction along that axis equals the current one.
    uint8_t last[8][8];
    // exact multi-GPU mode (k_sweep_columns_fused<.., LINK = true> only; see LinkSweep in sdfb_kernels.cuh)
    LinkSweep link;
    unsigned long long run_base;             // run << 32: link flag words are run_base | steps
    unsigned long long link_timeout_ns;      // watchdog of the waits on a neighbour GPU (0 = none)
    int cta_queue;                           // fused launches: 1 = column-wide evaluation queue, 0 = warp-private queues
    int link_gpu_fence;                      // SDFB_LINK_DEBUG & 2 (timing experiment): device-scope fence before the link flag
};

// smem exchange array: [2 slots][EK+1][EJ+1] words, index (b+1)*(EJ+1) + (a+1); a fastest
constexpr int RSTRIDE = (EK + 1) * (EJ + 1);
__device__ __forceinline__ int ring_idx(int a, int b) { return (b + 1) * (EJ + 1) + (a + 1); }

// named barriers: 0 = __syncthreads (column hand-over, all 352 threads), 1 = per-step (compute + halo lanes)
// The named barriers are warp-aligned and the compiler does not know that the inline asm is a convergent operation:
// the __syncwarp() guarantees that the warp is converged when it executes one, whatever divergent code came before.
__device__ __forceinline__ void bar_step() { __syncwarp(); asm volatile("bar.sync 1, %0;" ::"n"(NSTEPPERS) : "memory"); }
// 2 = evaluation barrier of the column-wide queue: compute AND halo lanes (the halo warps are nearly idle, so
// they take a share of the distance evaluations)
__device__ __forceinline__ void bar_compute() { __syncwarp(); asm volatile("bar.sync 2, %0;" ::"n"(NSTEPPERS) : "memory"); }

#ifdef SDFB_TRACE
#define TRACE(P, warp, s, slot) do { if ((P).trace && (threadIdx.x & 31) == 0 && sh.col == (P).trace_col) (P).trace[((size_t)(warp) * 8192 + (s)) * 8 + (slot)] = clock64(); } while (0)
#else
#define TRACE(P, warp, s, slot) do { } while (0)
#endif

// L2 prefetch of a run of cells of one row with ONE bulk request (TMA prefetch, no data returned).  The sweeps
// walk ~85,000 rows concurrently, 8 bytes per row and step; fetching them sector by sector makes every DRAM
// access a row miss.  `first` .. `first + (count-1)*dir` are relative positions along the row pointed to by p0.
constexpr int PF_CELLS = 64;          // cells per prefetch (512 B)
constexpr int PF_AHEAD = 96;          // how far ahead of the walker the prefetched run starts
__device__ __forceinline__ void prefetch_run(const uint64_t *p0, int64_t si, int first, int ni)
{
    int lo = first, hi = first + PF_CELLS - 1;                 // relative positions, clipped to the row
    if (hi > ni - 1) hi = ni - 1;
    if (lo < 0) lo = 0;
    if (lo > hi) return;
    const uint64_t *a = p0 + si * (int64_t)lo, *b = p0 + si * (int64_t)hi;
    const uint64_t *beg = si > 0 ? a : b;                        // lowest address of the run
    uintptr_t addr = reinterpret_cast<uintptr_t>(beg) & ~(uintptr_t)15;
    uint32_t bytes = (uint32_t)((hi - lo + 1) * 8 + 16) & ~15u;  // 16-byte granules covering the run
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(addr), "r"(bytes) : "memory");
}

struct ColShared {
    uint32_t ring[2 * RSTRIDE];
    uint32_t q_ent[NCOMPUTE / 32][QCAP];      // (owner lane << 27) | tri
    float q_d[NCOMPUTE / 32][QCAP];
    float gx[NCOMPUTE], gy[NCOMPUTE], gz[NCOMPUTE];   // world position of each compute lane's current voxel
    int wtot[NCOMPUTE / 32];                          // candidates per warp in this step (column-wide queue mode)
    int qn;                                           // SDFB_QATOMIC: entries reserved in this step (zeroed by lane 0 after use)
    int col;
    volatile int done;      // chunks whose steps are complete; written by compute lane 0
    float4 stage[STAGE > 0 ? NCOMPUTE * STAGE * 3 : 1];   // SDFB_STAGE: [owner lane][slot] -> the record's three float4
};

// Which entry an evaluating lane starts with.  The entries beyond a multiple of EVAL_LANES make a partial extra round for the
// lanes with the lowest start numbers.  SDFB_EVAL_REV=1 numbers the lanes backwards, so that this extra round falls on the
// halo warp (which has little else to do) and the LAST compute warps instead of on compute warp 0 -- the warp that shares
// its scheduler with the sync warp's polling loop when warps go to schedulers by warp number.
#ifndef SDFB_EVAL_REV
#define SDFB_EVAL_REV 0
#endif
__device__ __forceinline__ int eval_first(int lane_number) { return SDFB_EVAL_REV ? EVAL_LANES - 1 - lane_number : lane_number; }

// Evaluation share of one lane in the column-wide queue: entries first, first+EVAL_LANES, ...
__device__ __forceinline__ unsigned evaluate_queue_share(const TriRec *__restrict__ rec, ColShared &sh, int first, int total)
{
    uint32_t *const q_ent = &sh.q_ent[0][0];
    float *const q_d = &sh.q_d[0][0];
    unsigned evals = 0;
    int bnd[NCOMPUTE / 32];                       // SDFB_ONEBAR: first global entry number of each warp's region
    if (ONEBAR) {
        int acc = 0;
        #pragma unroll
        for (int w = 0; w < NCOMPUTE / 32; ++w) { bnd[w] = acc; acc += sh.wtot[w]; }
    }
    int qg = first;
    for (; qg < total; qg += EVAL_LANES) {
        int q = qg;
        if (ONEBAR) {                             // global entry number -> slot in the owning warp's region
            int w = 0, bw = 0;                    // selects, not bnd[w]: a dynamic index would put bnd[] in local memory
            #pragma unroll
            for (int u = 1; u < NCOMPUTE / 32; ++u) { const bool ge = qg >= bnd[u]; w += ge ? 1 : 0; bw = ge ? bnd[u] : bw; }
            q = w * QCAP + (qg - bw);
        }
#if SDFB_EVAL_PF && !SDFB_ONEBAR
        if (q + EVAL_LANES < total) {
            const char *ra = reinterpret_cast<const char *>(&rec[q_ent[q + EVAL_LANES]]);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(ra));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(ra + 32));
        }
#endif
        const int oe = __float_as_int(q_d[q]);                         // owner lane | staged slot << 8, replaced by the distance
        const int ot = oe & 0xff;
        const F3 gx{sh.gx[ot], sh.gy[ot], sh.gz[ot]};
        float4 p, qq, r;
        if (STAGE > 0 && (oe >> 8) < STAGE) {                          // the owner staged this record in shared memory
            const float4 *sr = &sh.stage[(ot * STAGE + (oe >> 8)) * 3];
            p = sr[0]; qq = sr[1]; r = sr[2];
        } else {
            const TriRec *tr = &rec[q_ent[q]];
            p = __ldg(&tr->p); qq = __ldg(&tr->q); r = __ldg(&tr->r);
        }
        q_d[q] = ptd_rec(gx, p, qq, r);
        ++evals;
    }
    return evals;
}

// What a halo lane does per step in the column-wide queue mode: the same three barriers as the compute
// lanes (one when the queue is empty) and its share of the evaluations.
__device__ __forceinline__ unsigned halo_evaluate_share(const TriRec *__restrict__ rec, ColShared &sh, int h)
{
    bar_compute();
    int total = 0;
    if (QATOMIC) {
        total = *reinterpret_cast<volatile int *>(&sh.qn);            // final: every warp reserved before the barrier
        if (total == 0) return 0u;
        const unsigned e = HALO_EVAL ? evaluate_queue_share(rec, sh, eval_first(NCOMPUTE + h), total) : 0u;
        bar_compute();
        return e;
    }
    #pragma unroll
    for (int w = 0; w < NCOMPUTE / 32; ++w) total += sh.wtot[w];
    if (total == 0) return 0u;
    if (!ONEBAR) bar_compute();                   // SDFB_ONEBAR: the entries were written before the first barrier
    const unsigned e = HALO_EVAL ? evaluate_queue_share(rec, sh, eval_first(NCOMPUTE + h), total) : 0u;
    bar_compute();
    return e;
}

// ---- sync warp: keeps flag polling and progress publication off the step loop's critical path ----------
// Lane 0 only.  Chunk c = steps [c*PUBLISH, (c+1)*PUBLISH).  A halo lane loads at step s the word of virtual
// step s+2, produced by column (J-1,K) lane (EJ-1,b) at its step s+2+EJ, so chunk c may run once the left
// column has completed s1-1+EJ+3 steps (EK for the column below; the diagonal column is covered transitively,
// because the left column itself waited for it).  Chunks are cleared a little ahead of need through sh.go;
// finished chunks (sh.done, set after the chunk's last step barrier) are fenced and published at once.
__device__ __forceinline__ void bar_go_arrive(int chunk) { __syncwarp(); asm volatile("bar.arrive %0, %1;" ::"r"(4 + (chunk & 3)), "n"(NHALO + 32) : "memory"); }
__device__ __forceinline__ void bar_go_wait(int chunk) { __syncwarp(); asm volatile("bar.sync %0, %1;" ::"r"(4 + (chunk & 3)), "n"(NHALO + 32) : "memory"); }

// Runs on the whole sync warp (lane 0 reads and writes; the named barriers need the full warp).  "Chunk c may
// run" is signalled to the halo warps with bar.arrive on barrier 4 + (c & 3): they block in hardware instead
// of spinning.  The sync warp clears at most 3 chunks beyond the finished ones, so when it re-arms a barrier
// (chunk c+4) its previous phase (chunk c) was consumed long ago.
// Fused launches: the compute lanes must not load their first cells before the previous sweep's columns are complete
// (in separate launches the cells are final when the kernel starts).  Named barrier 3: the sync warp arrives once the
// prerequisites hold, the compute lanes wait on it at the top of the column.
__device__ __forceinline__ void bar_start_arrive() { __syncwarp(); asm volatile("bar.arrive 3, %0;" ::"n"(NCOMPUTE + 32) : "memory"); }
__device__ __forceinline__ void bar_start_wait() { __syncwarp(); asm volatile("bar.sync 3, %0;" ::"n"(NCOMPUTE + 32) : "memory"); }

// Fused launches (several sweeps in one kernel, see k_sweep_columns_fused): before the column's first chunk is cleared,
// every column of the PREVIOUS sweep that intersects this column's rows grown by one row in all four directions must
// be complete (rule validated in oracle/experiments/sweep_overlap.c).  Q = the previous sweep's parameters, prevflags =
// its progress words.  Runs on the whole sync warp; lanes 0..8 poll one prerequisite column each.
__device__ __forceinline__ void wait_previous_sweep(const ColParams &P, const ColParams &Q, const uint32_t *prevflags,
                                                    int lane, int J, int K)
{
    const Grid &g = P.g;
    // this column's rows, relative then absolute, grown by one and clamped to the grid
    int rja = 1 + J * EJ, rjb = min(rja + EJ - 1, g.nj - 1);
    int rka = P.rk_first + K * EK, rkb = min(rka + EK - 1, P.rk_last);
    int j0 = P.sd.dj > 0 ? rja : g.nj - 1 - rjb, j1 = P.sd.dj > 0 ? rjb : g.nj - 1 - rja;
    int k0 = P.sd.dk > 0 ? rka : g.nk - 1 - rkb, k1 = P.sd.dk > 0 ? rkb : g.nk - 1 - rka;
    j0 = max(j0 - 1, 0); j1 = min(j1 + 1, g.nj - 1); k0 = max(k0 - 1, 0); k1 = min(k1 + 1, g.nk - 1);
    // -> rows of the previous sweep (relative), clipped to the rows it updates, -> its columns
    int qja = Q.sd.dj > 0 ? j0 : g.nj - 1 - j1, qjb = Q.sd.dj > 0 ? j1 : g.nj - 1 - j0;
    int qka = Q.sd.dk > 0 ? k0 : g.nk - 1 - k1, qkb = Q.sd.dk > 0 ? k1 : g.nk - 1 - k0;
    qja = max(qja, 1); qjb = min(qjb, g.nj - 1); qka = max(qka, Q.rk_first); qkb = min(qkb, Q.rk_last);
    if (qja > qjb || qka > qkb) return;                                   // uniform over the warp
    const int Ja = (qja - 1) / EJ, Jb = (qjb - 1) / EJ, Ka = (qka - Q.rk_first) / EK, Kb = (qkb - Q.rk_first) / EK;
    const int nJ = Jb - Ja + 1, n = nJ * (Kb - Ka + 1);                   // <= 3 x 3
    const uint32_t need = (Q.epoch << 16) + (uint32_t)Q.steps;
    const uint32_t *f = (lane < n) ? &prevflags[(Ka + lane / nJ) * Q.NJ + (Ja + lane % nJ)] : nullptr;
    for (;;) {
        const bool ok = !f || *reinterpret_cast<const volatile uint32_t *>(f) >= need;
        if (__all_sync(0xffffffffu, ok)) break;
        __nanosleep(SDFB_SYNC_SLEEP);
    }
    __threadfence();                    // acquire side: the column's loads (all ld.cg in fused launches) come after this
}

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// LINK (exact multi-GPU mode): `link_down` replaces prog_down for the columns of the FIRST K block when the slab below
// (in sweep direction) lives on another GPU -- same meaning (steps completed by the column that produces this column's
// b = -1 halo rows), 64-bit words run << 32 | steps written by that GPU with a system-scope release; `link_mine` is
// where the columns of the LAST K block publish for the slab above.  The words sit in the CONSUMER's memory, so
// polling them is a local L2 access.  A neighbour that never shows up (a host that died, mismatched calls) trips the
// watchdog: the kernel traps instead of spinning forever.
template <bool LINK>
__device__ __forceinline__ void sync_column(const ColParams &P, ColShared &sh, int lane, const uint32_t *prog_left,
                                            const uint32_t *prog_down, uint32_t *prog_mine,
                                            const unsigned long long *link_down = nullptr, unsigned long long *link_mine = nullptr)
{
    const uint32_t ebase = P.epoch << 16;
    const int nchunks = P.steps / PUBLISH;
    int cleared = 0, published = 0;                        // counts of chunks
    unsigned idle = 0;
    unsigned long long t_wait = 0;
    while (published < nchunks) {
        int act = 0, d = 0;                                // 1: clear the next chunk, 2: publish finished chunks
        if (lane == 0) {
            d = sh.done;
            if (cleared < nchunks && cleared < d + 3) {
                const int s1 = (cleared + 1) * PUBLISH;
                const uint32_t fl = prog_left ? *reinterpret_cast<const volatile uint32_t *>(prog_left) : 0xffffffffu;
                const uint32_t fd = prog_down ? *reinterpret_cast<const volatile uint32_t *>(prog_down) : 0xffffffffu;
                // no fence on this side: the halo lanes' loads are issued only after the barrier below (control
                // dependence) and go to L2 (ld.cg), where the producer's stores landed before its flag
                bool ok = fl >= ebase + (uint32_t)min(P.steps, s1 - 1 + EJ + 3) && fd >= ebase + (uint32_t)min(P.steps, s1 - 1 + EK + 3);
                if (LINK && link_down)
                    ok = ok && *reinterpret_cast<const volatile unsigned long long *>(link_down) >= P.run_base + (unsigned long long)min(P.steps, s1 - 1 + EK + 3);
                if (ok) act = 1;
            }
            if (!act && d > published) act = 2;
        }
        act = __shfl_sync(0xffffffffu, act, 0);
        d = __shfl_sync(0xffffffffu, d, 0);
        if (act == 1) {
            bar_go_arrive(cleared);
            ++cleared;
            idle = 0;
        } else if (act == 2) {
            if (lane == 0) {
                // RELEASE stores, not __threadfence() + store: the chunk's cell stores (made by the compute lanes, ordered
                // before sh.done by the CTA barrier and observed here through it) happen-before the flag, and that is all a
                // producer needs.  __threadfence() is fence.sc and ptxas follows it with CCTL.IVALL, which throws away the
                // SM's whole L1 -- three CTAs per SM doing that every two steps kept the triangle records (the only data
                // this kernel reads through L1) out of it.  st.release = MEMBAR.ALL + store, no invalidate.
                if (LINK && link_mine) {
                    // system scope: the boundary-plane cells stored into the neighbour's memory (peer stores over NVLink)
                    // are visible there before the word that announces them
                    const unsigned long long v = P.run_base + (unsigned long long)(d * PUBLISH);
                    if (P.link_gpu_fence) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(link_mine), "l"(v) : "memory");
                    else asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(link_mine), "l"(v) : "memory");
                }
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(prog_mine), "r"(ebase + (uint32_t)(d * PUBLISH)) : "memory");
            }
            published = d;
            idle = 0;
        } else {
            __nanosleep(SDFB_SYNC_SLEEP);
            if (LINK && P.link_timeout_ns && (++idle & 1023u) == 0) {            // uniform over the warp
                const unsigned long long now = global_timer_ns();
                if (idle == 1024u) t_wait = now;
                else if (now - t_wait > P.link_timeout_ns) {
                    if (lane == 0) printf("sdfb: column wait timed out (sweep stamp %u, waiting for %s)\n", P.stamp, link_down ? "the neighbour GPU" : "a local column");
                    __trap();
                }
            }
        }
    }
}

// ---- halo warps: feed the words of the upstream columns / boundary faces into the exchange array ----
// LINK (exact multi-GPU mode): the b = -1 rows of a column of the FIRST K block lie in the slab below; when that slab
// lives on another GPU they are read from this sweep's inbound plane (P.link.halo_src, filled by the neighbour's last
// K block while it runs the same sweep) instead of the halo plane of the cell array.  And the one row no column of
// the neighbour ever computes -- rj = 0, read-only in this sweep -- is forwarded to the slab above by the a = -1 halo
// lane of the J = 0 column that reads it anyway (b = the slab's last plane).
template <bool CTA_QUEUE, bool LINK>
__device__ __forceinline__ void halo_column(const uint64_t *__restrict__ cells, const TriRec *__restrict__ rec,
                                            const ColParams &P, ColShared &sh, int h, int rj0, int rk0, unsigned &my_evals,
                                            const uint32_t *prog_left = nullptr, const uint32_t *prog_down = nullptr, uint32_t *prog_mine = nullptr,
                                            const unsigned long long *link_down = nullptr, unsigned long long *link_mine = nullptr)
{
    // SDFB_MERGE_SYNC: what this warp last read from its producers' progress words (all lanes hold the same values)
    const uint32_t ebase = P.epoch << 16;
    uint32_t seen_l = prog_left ? 0u : 0xffffffffu, seen_d = prog_down ? 0u : 0xffffffffu;
    unsigned long long seen_link = (LINK && link_down) ? 0ull : ~0ull;
    const Grid &g = P.g;
    int a, b;
    if (h <= EK) { a = -1; b = h - 1; }                    // (-1,-1), (-1,0) .. (-1,EK-1)
    else if (h <= EK + EJ) { a = h - EK - 1; b = -1; }     // (0,-1) .. (EJ-1,-1)
    else { a = -2; b = -2; }                               // idle lanes
    const int rj = rj0 + a, rk = rk0 + b;
    const bool row_ok = (a > -2) && rj <= g.nj - 1 && rk <= P.rk_last;
    const int64_t si = (int64_t)P.sd.di;
    const uint64_t *ptr = cells;
    if (row_ok) ptr = cells + g.cidx(P.sd.abs_i(0, g), P.sd.abs_j(rj, g), P.sd.abs_k(rk, g)) + si * (int64_t)(0 - a - b - SHIFT);
    uint64_t *fwd = nullptr;                               // LINK: where this lane forwards row rj = 0 of the boundary plane
    if (LINK && row_ok) {
        if (P.link.halo_src && b == -1 && rk0 == P.rk_first)
            ptr = P.link.halo_src + ((int64_t)P.sd.abs_i(0, g) + (int64_t)g.ni * P.sd.abs_j(rj, g)) + si * (int64_t)(0 - a - b - SHIFT);
        if (P.link.halo_dst && a == -1 && rj == 0 && rk == P.rk_last)
            fwd = P.link.halo_dst + ((int64_t)P.sd.abs_i(0, g) + (int64_t)g.ni * P.sd.abs_j(0, g)) + si * (int64_t)(0 - a - b - SHIFT);
    }
    const int widx = row_ok ? ring_idx(a, b) : 0;
    int ri = 0 - a - b - SHIFT;                            // voxel of virtual step 0
    // Raw cells of virtual steps s (even -> wA, odd -> wB), each loaded two steps before it is published.
    // The step loop is unrolled by two so that a register is reloaded right after it was consumed and
    // never copied: a copy (or a select) of a freshly loaded value would stall on the load at once.
    uint64_t wA = ~0ull, wB = ~0ull;
    for (int s0 = 0, c = 0; s0 < P.steps; s0 += PUBLISH, ++c) {
        const int s1 = s0 + PUBLISH;                       // P.steps is a multiple of PUBLISH
        if (!MERGE) bar_go_wait(c);                        // until the sync warp has cleared chunk c
        else {
            // the same condition as sync_column's, polled here: the loads of this chunk (words of steps up to s1 + 1) may be
            // issued once the producers have completed s1 - 1 + E + 3 steps.  Uniform over the warp; no fence (see sync_column).
            const uint32_t need_l = ebase + (uint32_t)min(P.steps, s1 - 1 + EJ + 3), need_d = ebase + (uint32_t)min(P.steps, s1 - 1 + EK + 3);
            const unsigned long long need_link = P.run_base + (unsigned long long)min(P.steps, s1 - 1 + EK + 3);
            unsigned idle = 0;
            unsigned long long t_wait = 0;
            while (seen_l < need_l || seen_d < need_d || (LINK && seen_link < need_link)) {
                if (prog_left) seen_l = *reinterpret_cast<const volatile uint32_t *>(prog_left);
                if (prog_down) seen_d = *reinterpret_cast<const volatile uint32_t *>(prog_down);
                if (LINK && link_down) seen_link = *reinterpret_cast<const volatile unsigned long long *>(link_down);
                if (seen_l >= need_l && seen_d >= need_d && !(LINK && seen_link < need_link)) break;
                __nanosleep(SDFB_SYNC_SLEEP);
                if (LINK && P.link_timeout_ns && (++idle & 1023u) == 0) {
                    const unsigned long long now = global_timer_ns();
                    if (idle == 1024u) t_wait = now;
                    else if (now - t_wait > P.link_timeout_ns) {
                        if (h == 0) printf("sdfb: column wait timed out (sweep stamp %u, waiting for %s)\n", P.stamp, link_down ? "the neighbour GPU" : "a local column");
                        __trap();
                    }
                }
            }
        }
        if (s0 == 0 && row_ok) {
            if ((unsigned)ri < (unsigned)g.ni) wA = __ldcg(ptr);
            if ((unsigned)(ri + 1) < (unsigned)g.ni) wB = __ldcg(ptr + si);
        }
        for (int s = s0; s < s1; s += 2) {                 // PUBLISH is even
            TRACE(P, 8 + (h >> 5), s, 0);
            if (row_ok) {
                sh.ring[widx] = cell_lo(wA);               // even step -> slot 0
                if (LINK && fwd && (unsigned)ri < (unsigned)g.ni) *fwd = wA;
                wA = ~0ull;
                if ((unsigned)(ri + 2) < (unsigned)g.ni && s + 2 < P.steps) wA = __ldcg(ptr + 2 * si);
            }
            if (CTA_QUEUE) my_evals += halo_evaluate_share(rec, sh, h);
            TRACE(P, 8 + (h >> 5), s, 7);
            bar_step();
            TRACE(P, 8 + (h >> 5), s + 1, 0);
            if (row_ok) {
                sh.ring[RSTRIDE + widx] = cell_lo(wB);     // odd step -> slot 1
                if (LINK && fwd) { if ((unsigned)(ri + 1) < (unsigned)g.ni) *(fwd + si) = wB; fwd += 2 * si; }
                wB = ~0ull;
                if ((unsigned)(ri + 3) < (unsigned)g.ni && s + 3 < P.steps) wB = __ldcg(ptr + 3 * si);
                ri += 2; ptr += 2 * si;
                // non-binding L2 prefetch far ahead (the line may still be rewritten by its producer; L2 stays coherent)
                if ((s & (PF_CELLS - 1)) == 0) prefetch_run(ptr - si * (int64_t)ri, si, ri + PF_AHEAD, g.ni);
            }
            if (CTA_QUEUE) my_evals += halo_evaluate_share(rec, sh, h);
            TRACE(P, 8 + (h >> 5), s + 1, 7);
            bar_step();
        }
        if (MERGE && h == 0) {
            // every lane's cell stores of this chunk precede its arrival at the step barrier this warp has just left: the
            // RELEASE store (not fence + store, see sync_column) makes them visible before the progress word
            if (LINK && link_mine) {
                const unsigned long long v = P.run_base + (unsigned long long)s1;
                if (P.link_gpu_fence) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(link_mine), "l"(v) : "memory");
                else asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(link_mine), "l"(v) : "memory");
            }
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(prog_mine), "r"(ebase + (uint32_t)s1) : "memory");
        }
    }
}

// ---- compute warps ---------------------------------------------------------------------------------
// Per-lane state of a compute lane, kept in registers across the (unrolled) step loop.
struct LaneState {
    uint64_t *own_ptr;                    // cell of the current step's voxel
    int ri;                               // its relative i
    uint32_t prev_lo;                     // own result of step s-1                                  -> m=0
    // words read from the exchange array in earlier steps, rolled through registers:
    //   R1(s) = lane(a-1,b  )@s-1 = (ri,   rj-1, rk  )   m=1 now, m=2 one step later
    //   R3(s) = lane(a,  b-1)@s-1 = (ri,   rj,   rk-1)   m=3 now, m=4 one step later
    //   R5(s) = lane(a-1,b-1)@s-1 = (ri+1, rj-1, rk-1)   m=5 one step later, m=6 two steps later
    uint32_t r1_old, r3_old, r5_old, r5_old2;
    unsigned changed, evals;
    uint64_t *push_ptr;                   // LINK: this voxel's cell in the inbound plane of the slab above (nullptr: not a boundary lane)
};

// The rare part of a step: at least one lane of the warp has a neighbour whose cell changed since the
// lane last looked at that offset.  Filters the candidates, balances the distance evaluations over the
// warp through a queue in shared memory and replays each lane's results in the reference's order.
// Everything is passed and returned by value (registers): a reference to the caller's arrays would force
// them into local memory on the hot path.  Returns {new cell word, (evaluations << 1) | changed}.
__device__ __forceinline__ uint2 evaluate_candidates(const TriRec *__restrict__ rec, const ColParams &P, ColShared &sh,
                                                  int ri, int lane, int warp, bool update, bool edge,
                                                  uint32_t nb0, uint32_t nb1, uint32_t nb2, uint32_t nb3, uint32_t nb4,
                                                  uint32_t nb5, uint32_t nb6, uint32_t th0, uint32_t th1, uint32_t th2,
                                                  uint32_t th3, uint32_t th4, uint32_t th5, uint32_t th6,
                                                  uint32_t cur, uint64_t *self_ptr, float phi, float &phi_new)
{
    const Grid &g = P.g;
    uint32_t *const q_ent = sh.q_ent[warp];
    float *const q_d = sh.q_d[warp];
    const uint32_t nb[7] = {nb0, nb1, nb2, nb3, nb4, nb5, nb6}, thr[7] = {th0, th1, th2, th3, th4, th5, th6};
    unsigned evals = 0, changed = 0;
    uint32_t live = 0;                    // bit m: neighbour m's triangle must be evaluated
    if (update) {
        // keep m if it names a triangle, not the voxel's own, and (memo) its cell changed since this voxel
        // last looked at offset m
        #pragma unroll
        for (int m = 0; m < 7; ++m) {
            const uint32_t x = nb[m];
            const bool keep = ((x & TRI_MASK) != TRI_NONE) && (((x ^ cur) & TRI_MASK) != 0) && (edge || x >= thr[m]);
            live |= keep ? (1u << m) : 0u;
        }
        if (live) {                       // drop repeats of ANY earlier neighbour's triangle: that triangle is
            #pragma unroll                // either evaluated there, or the voxel's own, or a known loser (memo)
            for (int m = 1; m < 7; ++m) {
                bool dup = false;
                #pragma unroll
                for (int u = 0; u < m; ++u) dup = dup || (((nb[u] ^ nb[m]) & TRI_MASK) == 0);
                if (dup) live &= ~(1u << m);
            }
        }
    }
    const int ncand = __popc(live);
    // exclusive scan of the candidate counts over the warp (3 ballots: ncand <= 7)
    const uint32_t b0 = __ballot_sync(0xffffffffu, ncand & 1), b1 = __ballot_sync(0xffffffffu, ncand & 2),
                   b2 = __ballot_sync(0xffffffffu, ncand & 4);
    if ((b0 | b1 | b2) == 0) return make_uint2(cur, 0u);
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int total = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
    const int off = __popc(b0 & lt_mask) + 2 * __popc(b1 & lt_mask) + 4 * __popc(b2 & lt_mask);
    if (live) {
        sh.gx[(warp << 5) + lane] = lattice(P.sd.abs_i(ri, g), g.dx, g.ox);      // gy, gz were stored once per column
        int w = off;
        #pragma unroll
        for (int m = 0; m < 7; ++m) if ((live >> m) & 1u) {
            q_ent[w] = ((uint32_t)lane << 27) | (nb[m] & TRI_MASK); ++w;
            const char *ra = reinterpret_cast<const char *>(&rec[nb[m] & TRI_MASK]);     // start the gather now
            asm volatile("prefetch.global.L1 [%0];" ::"l"(ra));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(ra + 32));
        }
    }
    __syncwarp();
    for (int q = lane; q < total; q += 32) {
        const uint32_t e = q_ent[q];
        const int ot = (warp << 5) + (int)(e >> 27);                 // owner lane -> its voxel's position
        const F3 gx{sh.gx[ot], sh.gy[ot], sh.gz[ot]};
        const TriRec *tr = &rec[e & TRI_MASK];
        const float4 p = __ldg(&tr->p), qq = __ldg(&tr->q), r = __ldg(&tr->r);
        q_d[q] = ptd_rec(gx, p, qq, r);
        ++evals;
    }
    __syncwarp();
    if (live) {
        uint32_t best = TRI_NONE;
        for (int q = off; q < off + ncand; ++q) {                    // the reference's order and strict "<"
            const float d = q_d[q];
            if (d < phi) { phi = d; best = q_ent[q] & TRI_MASK; }
        }
        if (best != TRI_NONE) {
            cur = (P.stamp << 27) | best;
            *self_ptr = pack_cell(phi, cur);
            phi_new = phi;
            changed = 1;
        }
    }
    __syncwarp();
    return make_uint2(cur, (evals << 1) | changed);
}

// Column-wide variant used in the evaluation-heavy sweeps: the candidates of all 8 compute warps go to one
// queue and all 256 lanes evaluate it, so every warp does ceil(total/256) rounds instead of waiting at the
// step barrier for the warp with the longest private queue.  Costs one compute-only barrier per step (three
// when there is work).  Must be called by every compute lane in every step.
__device__ __forceinline__ uint2 evaluate_candidates_cta(const TriRec *__restrict__ rec, const ColParams &P, ColShared &sh,
                                                      int ri, int tid, int s, bool update, uint32_t live_in,
                                                      uint32_t nb0, uint32_t nb1, uint32_t nb2, uint32_t nb3, uint32_t nb4,
                                                      uint32_t nb5, uint32_t nb6, uint32_t cur, uint64_t *self_ptr, float phi,
                                                      float &phi_new)
{
    const Grid &g = P.g;
    const int lane = tid & 31, warp = tid >> 5;
    phi_new = phi;                                 // LINK: the caller forwards the voxel's final cell to the slab above
    uint32_t *const q_ent = &sh.q_ent[0][0];       // flat: NCOMPUTE * 7 entries
    float *const q_d = &sh.q_d[0][0];
    const uint32_t nb[7] = {nb0, nb1, nb2, nb3, nb4, nb5, nb6};
    unsigned evals = 0, changed = 0;
    uint32_t live = update ? live_in : 0u;
    if (live) {                           // drop repeats of ANY earlier neighbour's triangle (see evaluate_candidates)
        #pragma unroll
        for (int m = 1; m < 7; ++m) {
            bool dup = false;
            #pragma unroll
            for (int u = 0; u < m; ++u) dup = dup || (((nb[u] ^ nb[m]) & TRI_MASK) == 0);
            if (dup) live &= ~(1u << m);
        }
    }
    const int ncand = __popc(live);
    if (STAGE > 0 && live) {
        // start the gather of the first STAGE candidates' records into this lane's slots (16-byte cp.async, L1-allocating);
        // completion is awaited before the second queue barrier, which publishes the slots to the evaluating lanes
        int slot = 0;
        #pragma unroll
        for (int m = 0; m < 7; ++m) if (((live >> m) & 1u) && slot < STAGE) {
            const char *src = reinterpret_cast<const char *>(&rec[nb[m] & TRI_MASK]);
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&sh.stage[(tid * STAGE + slot) * 3]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16), "l"(src + 16) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 32), "l"(src + 32) : "memory");
            ++slot;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const uint32_t b0 = __ballot_sync(0xffffffffu, ncand & 1), b1 = __ballot_sync(0xffffffffu, ncand & 2),
                   b2 = __ballot_sync(0xffffffffu, ncand & 4);
    const uint32_t lt_mask = (1u << lane) - 1u;
    int base = 0, total = 0;
    if (QATOMIC) {
        const int wt = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
        if (wt) {                                                     // uniform over the warp
            if (lane == 0) base = atomicAdd(&sh.qn, wt);
            base = __shfl_sync(0xffffffffu, base, 0);
        }
    } else if (ONEBAR) {
        if (lane == 0) sh.wtot[warp] = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
        base = warp * QCAP;                                           // this warp's own region: no other warp's total needed
    } else {
        if (lane == 0) sh.wtot[warp] = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
        TRACE(P, warp, s, 1);
        bar_compute();
        TRACE(P, warp, s, 2);
        #pragma unroll
        for (int w = 0; w < NCOMPUTE / 32; ++w) { const int t = sh.wtot[w]; base += (w < warp) ? t : 0; total += t; }
        if (total == 0) return make_uint2(cur, 0u);                   // uniform over the column
    }
    const int off = base + __popc(b0 & lt_mask) + 2 * __popc(b1 & lt_mask) + 4 * __popc(b2 & lt_mask);
    if (live) {
        sh.gx[tid] = lattice(P.sd.abs_i(ri, g), g.dx, g.ox);           // gy, gz were stored once per column
        int w = off;
        #pragma unroll
        for (int m = 0; m < 7; ++m) if ((live >> m) & 1u) {
            q_ent[w] = nb[m] & TRI_MASK;
            q_d[w] = __int_as_float(tid | ((w - off) << 8));           // owner | slot, replaced by the distance below
            ++w;
#if SDFB_ENQ_PF
            const char *ra = reinterpret_cast<const char *>(&rec[nb[m] & TRI_MASK]);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(ra));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(ra + 32));
#endif
        }
    }
    if (STAGE > 0) asm volatile("cp.async.wait_all;" ::: "memory");  // this lane's staged records have landed (no-op without any)
    TRACE(P, warp, s, 3);
    bar_compute();
    TRACE(P, warp, s, 4);
    if (QATOMIC) {
        total = *reinterpret_cast<volatile int *>(&sh.qn);            // final: every warp reserved before the barrier
        if (total == 0) return make_uint2(cur, 0u);                   // uniform over the column
    }
    if (ONEBAR) {
        #pragma unroll
        for (int w = 0; w < NCOMPUTE / 32; ++w) total += sh.wtot[w];
        if (total == 0) return make_uint2(cur, 0u);                   // uniform over the column
    }
    evals = evaluate_queue_share(rec, sh, eval_first(tid), total);
    TRACE(P, warp, s, 5);
    bar_compute();
    TRACE(P, warp, s, 6);
    if (QATOMIC && tid == 0) sh.qn = 0;     // every lane read the total before this barrier; the next reservation comes after bar_step
    if (live) {
        // the reference's order and strict "<": all distances are fetched first (independent loads)
        float dv[7];
        #pragma unroll
        for (int m = 0; m < 7; ++m) dv[m] = (m < ncand) ? q_d[off + m] : __int_as_float(0x7f800000);
        int best = -1;
        #pragma unroll
        for (int m = 0; m < 7; ++m) if (dv[m] < phi) { phi = dv[m]; best = m; }
        if (best >= 0) {
            cur = (P.stamp << 27) | q_ent[off + best];
            *self_ptr = pack_cell(phi, cur);
            phi_new = phi;
            changed = 1;
        }
    }
    return make_uint2(cur, (evals << 1) | changed);
}

// One step of a compute lane.  PAR = step parity: reads exchange slot PAR^1, writes slot PAR.  `own` holds the
// lane's cell for this step on entry and is reloaded with the cell two steps ahead (see halo_column).
template <int PAR, bool CTA_QUEUE, bool L2OWN = false, bool LINK = false>
__device__ __forceinline__ void compute_step(const TriRec *__restrict__ rec, const ColParams &P, ColShared &sh,
                                             const uint32_t *ring_r, uint32_t *ring_w, int s, int lane, int warp,
                                             int rj0, int rk0, bool row_ok, const uint32_t (&thr)[7],
                                             const uint32_t (&thr_edge)[7], uint64_t &own, const uint64_t &own_other,
                                             LaneState &st)
{
    if (CTA_QUEUE) {
        // evaluation-heavy sweeps: the triangle of the lane's NEXT voxel (cell loaded one step ago) will be a
        // candidate of its neighbours one or two steps from now -- start the record gather a step early
        const uint32_t tn = lo_tri(cell_lo(own_other));
        if (row_ok && tn != TRI_NONE) {
            const char *ra = reinterpret_cast<const char *>(&rec[tn]);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ra));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ra + 32));
        }
    }
    const int ni = P.g.ni;
    const int64_t si = (int64_t)P.sd.di;
    const int ri = st.ri;
    TRACE(P, warp, s, 0);
    const uint32_t *rr = ring_r + (PAR ^ 1) * RSTRIDE;
    const uint32_t r5 = rr[0], r3 = rr[1], r1 = rr[EJ + 1];
    const uint64_t self = own;
    uint64_t *const self_ptr = st.own_ptr;
    // the cell two steps ahead.  Fused launches: another SM wrote it in the previous sweep of the SAME launch, so this
    // SM's L1 may hold a stale line from the sweep before that -> read it from L2
    if (row_ok && (unsigned)(ri + 2) < (unsigned)ni) own = L2OWN ? __ldcg(self_ptr + 2 * si) : *(self_ptr + 2 * si);
    if (row_ok && (s & (PF_CELLS - 1)) == 0) prefetch_run(self_ptr - si * (int64_t)ri, si, ri + PF_AHEAD, ni);
    uint32_t cur = cell_lo(self);
    const bool update = row_ok && (unsigned)(ri - 1) < (unsigned)(ni - 1);            // 1 <= ri <= ni-1
    // The last voxel of a row lies on the far i face: only sweeps with the same di visit it (thr_edge).
    const bool edge = (ri == ni - 1);
    const uint32_t nb[7] = {st.prev_lo, r1, st.r1_old, r3, st.r3_old, st.r5_old, st.r5_old2};
    uint32_t t[7];
    bool fresh = false;
    #pragma unroll
    for (int m = 0; m < 7; ++m) { t[m] = edge ? thr_edge[m] : thr[m]; fresh = fresh || (nb[m] >= t[m]); }
    float phi_new = cell_phi(self);
    if (CTA_QUEUE) {
        // keep m if it names a triangle, not the voxel's own, and (memo) its cell changed since this voxel last looked
        uint32_t live = 0;
        if (update && fresh) {
            #pragma unroll
            for (int m = 0; m < 7; ++m) {
                const uint32_t x = nb[m];
                const bool keep = ((x & TRI_MASK) != TRI_NONE) && (((x ^ cur) & TRI_MASK) != 0) && (x >= t[m]);
                live |= keep ? (1u << m) : 0u;
            }
        }
        const uint2 r = evaluate_candidates_cta(rec, P, sh, ri, (warp << 5) + lane, s, update, live,
                                                nb[0], nb[1], nb[2], nb[3], nb[4], nb[5], nb[6], cur, self_ptr, cell_phi(self), phi_new);
        cur = r.x; st.changed += r.y & 1u; st.evals += r.y >> 1;
    } else if (__any_sync(0xffffffffu, update && fresh)) {
        const uint2 r = evaluate_candidates(rec, P, sh, ri, lane, warp, update, false,
                                            nb[0], nb[1], nb[2], nb[3], nb[4], nb[5], nb[6],
                                            t[0], t[1], t[2], t[3], t[4], t[5], t[6],
                                            cur, self_ptr, cell_phi(self), phi_new);
        cur = r.x; st.changed += r.y & 1u; st.evals += r.y >> 1;
    }
    // LINK: a lane on the slab's last plane stores the voxel's final cell (changed or not, ri = 0 included) into the
    // inbound plane of the slab above; the sync warp publishes the step count there after a system-scope fence
    if (LINK && st.push_ptr) {
        if ((unsigned)ri < (unsigned)ni) *st.push_ptr = pack_cell(phi_new, cur);
        st.push_ptr += si;
    }
    st.r1_old = r1; st.r3_old = r3; st.r5_old2 = st.r5_old; st.r5_old = r5;
    if (row_ok && (unsigned)ri < (unsigned)ni) { ring_w[PAR * RSTRIDE] = cur; st.prev_lo = cur; }
    st.own_ptr = self_ptr + si;
    st.ri = ri + 1;
    TRACE(P, warp, s, 7);
    bar_step();
}

template <bool CTA_QUEUE, bool L2OWN = false, bool LINK = false>
__device__ __forceinline__ void compute_column(uint64_t *__restrict__ cells, const TriRec *__restrict__ rec,
                                               const ColParams &P, ColShared &sh, int tid, int rj0, int rk0,
                                               unsigned &my_changed, unsigned &my_evals)
{
    const Grid &g = P.g;
    const int lane = tid & 31, warp = tid >> 5;
    const int a = tid % EJ, b = tid / EJ;
    const int rj = rj0 + a, rk = rk0 + b;
    const bool row_ok = rj <= g.nj - 1 && rk <= P.rk_last;
    const int64_t si = (int64_t)P.sd.di;
    LaneState st;
    st.own_ptr = cells;
    st.push_ptr = nullptr;
    st.ri = 0 - a - b - SHIFT;            // voxel of step 0
    if (row_ok) {
        const int j = P.sd.abs_j(rj, g), k = P.sd.abs_k(rk, g);
        st.own_ptr = cells + g.cidx(P.sd.abs_i(0, g), j, k) + si * (int64_t)st.ri;
        sh.gy[tid] = lattice(j, g.dx, g.oy);
        sh.gz[tid] = lattice(k, g.dx, g.oz);
        if (LINK && P.link.halo_dst && rk == P.rk_last)
            st.push_ptr = P.link.halo_dst + ((int64_t)P.sd.abs_i(0, g) + (int64_t)g.ni * j) + si * (int64_t)st.ri;
    }
    // memo thresholds: the neighbour word nb at offset m is fresh iff nb >= thr[m] = (last[m]+1) << 27, i.e.
    // stamp(nb) > last[m] (0 = always fresh where the offset was never examined).  Rows on the far j / k
    // face use the table of their class; the last voxel of a row (far i face) has its own thresholds.
    const int row_class = (rj == g.nj - 1 ? 2 : 0) | (rk == g.nk - 1 ? 4 : 0);
    uint32_t thr[7], thr_edge[7];
    #pragma unroll
    for (int m = 0; m < 7; ++m) {
        const uint32_t l0 = P.last[row_class][m], l1 = P.last[row_class | 1][m];
        thr[m] = l0 ? (l0 + 1u) << 27 : 0u;
        thr_edge[m] = l1 ? (l1 + 1u) << 27 : 0u;
    }
    const uint32_t *ring_r = sh.ring + ring_idx(a - 1, b - 1);     // R5 at +0, R3 at +1, R1 at +(EJ+1)
    uint32_t *ring_w = sh.ring + ring_idx(a, b);
    st.prev_lo = TRI_NONE; st.r1_old = TRI_NONE; st.r3_old = TRI_NONE; st.r5_old = TRI_NONE; st.r5_old2 = TRI_NONE;
    st.changed = 0; st.evals = 0;
    // own cells by step parity (even -> ownA, odd -> ownB), each loaded two steps before use and reloaded
    // right after it was consumed (no copies of fresh loads, see halo_column)
    uint64_t ownA = 0, ownB = 0;
    if (row_ok) {                          // the first stretch of the row, before the periodic prefetch takes over
        prefetch_run(st.own_ptr - si * (int64_t)st.ri, si, 0, g.ni);
        prefetch_run(st.own_ptr - si * (int64_t)st.ri, si, PF_CELLS, g.ni);
    }
    if (L2OWN) bar_start_wait();           // fused launches: the previous sweep is complete on the rows this column touches
    if (row_ok && (unsigned)st.ri < (unsigned)g.ni) ownA = L2OWN ? __ldcg(st.own_ptr) : *st.own_ptr;
    if (row_ok && (unsigned)(st.ri + 1) < (unsigned)g.ni) ownB = L2OWN ? __ldcg(st.own_ptr + si) : *(st.own_ptr + si);
    for (int s = 0; s < P.steps; s += 2) {   // P.steps is even
        compute_step<0, CTA_QUEUE, L2OWN, LINK>(rec, P, sh, ring_r, ring_w, s, lane, warp, rj0, rk0, row_ok, thr, thr_edge, ownA, ownB, st);
        compute_step<1, CTA_QUEUE, L2OWN, LINK>(rec, P, sh, ring_r, ring_w, s + 1, lane, warp, rj0, rk0, row_ok, thr, thr_edge, ownB, ownA, st);
        if (tid == 0 && ((s + 2) % PUBLISH) == 0) {      // every lane's stores of this chunk precede the barrier
            __threadfence_block();
            sh.done = (s + 2) / PUBLISH;
        }
    }
    my_changed += st.changed; my_evals += st.evals;
}

// MINB = CTAs per SM the register allocation is bounded for.  3 (93 registers, no spills) is slightly faster where the
// wavefront is narrow and the kernel latency-bound (512^3: 60.5 vs 61.4 ms for the first pass); 4 (80 registers, a few
// spills) wins where there is enough work to be throughput-bound (1024^3: 240 vs 263 ms).
// The column loop of one CTA.  GROUP selects the roles compiled in: -1 = all (one register budget for the whole CTA),
// 0 = the compute warps, 1 = the halo, sync and idle warps (SDFB_WG: each warpgroup runs its own copy of the loop in
// the branch its setmaxnreg dominates, which is what makes ptxas allocate registers per role).  Every thread of the
// CTA passes the same two __syncthreads() per column whichever copy it runs.
template <bool CTA_QUEUE, int GROUP>
__device__ __forceinline__ void column_loop(uint64_t *__restrict__ cells, const TriRec *__restrict__ rec, const ColParams &P,
                                            ColShared &sh, uint32_t *__restrict__ progress, uint32_t *__restrict__ ticket,
                                            int tid, int lane, int ncols, unsigned &my_changed, unsigned &my_evals)
{
    for (;;) {
        // ---- take the next column (anti-diagonal order) --------------------------------------
        if (GROUP != 1 && tid == 0) { sh.col = (int)atomicAdd(ticket, 1u); sh.done = 0; sh.qn = 0; }
        __syncthreads();
        const int tk = sh.col;
        if (tk >= ncols) break;
        int J, K;
        {
            int d = 0, rem = tk;
            for (;;) {
                int lo = max(0, d - (P.NK - 1)), hi = min(d, P.NJ - 1);
                int cnt = hi - lo + 1;
                if (rem < cnt) { J = lo + rem; K = d - J; break; }
                rem -= cnt; ++d;
            }
        }
        const int rj0 = 1 + J * EJ, rk0 = P.rk_first + K * EK;
        if (GROUP != 1 && tid < NCOMPUTE) {
            compute_column<CTA_QUEUE>(cells, rec, P, sh, tid, rj0, rk0, my_changed, my_evals);
        } else if (GROUP != 0 && tid >= NCOMPUTE && tid < NSTEPPERS) {
            if (MERGE) {
                const uint32_t *prog_left = (J > 0) ? &progress[K * P.NJ + (J - 1)] : nullptr;
                const uint32_t *prog_down = (K > 0) ? &progress[(K - 1) * P.NJ + J] : nullptr;
                halo_column<CTA_QUEUE, false>(cells, rec, P, sh, tid - NCOMPUTE, rj0, rk0, my_evals, prog_left, prog_down, &progress[K * P.NJ + J]);
            } else {
                halo_column<CTA_QUEUE, false>(cells, rec, P, sh, tid - NCOMPUTE, rj0, rk0, my_evals);
            }
        } else if (!MERGE && GROUP != 0 && tid >= NSTEPPERS && tid < NSTEPPERS + 32) {
            const uint32_t *prog_left = (J > 0) ? &progress[K * P.NJ + (J - 1)] : nullptr;
            const uint32_t *prog_down = (K > 0) ? &progress[(K - 1) * P.NJ + J] : nullptr;
            sync_column<false>(P, sh, lane, prog_left, prog_down, &progress[K * P.NJ + J]);
        }
        __syncthreads();        // sh.col is rewritten next; also orders the two roles' exits
    }
}

template <bool CTA_QUEUE, int MINB>
__global__ void __launch_bounds__(NTHREADS, WG ? 4 : minb_bound(MINB))
k_sweep_columns(uint64_t *__restrict__ cells, const TriRec *__restrict__ rec, ColParams P,
                uint32_t *__restrict__ progress, uint32_t *__restrict__ ticket,
                unsigned long long *__restrict__ changed)
{
    __shared__ ColShared sh;
    const int tid = threadIdx.x, lane = tid & 31;
    const int ncols = P.NJ * P.NK;
    unsigned my_changed = 0, my_evals = 0;
    if (P.run_if && __ldcg(P.run_if) == 0u) return;                   // uniform over the grid; the ticket is untouched
    if (WG) {
        if (tid < NCOMPUTE) {
            asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(SDFB_WG_COMPUTE_REGS));
            column_loop<CTA_QUEUE, 0>(cells, rec, P, sh, progress, ticket, tid, lane, ncols, my_changed, my_evals);
        } else {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SDFB_WG_LIGHT_REGS));
            column_loop<CTA_QUEUE, 1>(cells, rec, P, sh, progress, ticket, tid, lane, ncols, my_changed, my_evals);
        }
    } else {
        column_loop<CTA_QUEUE, -1>(cells, rec, P, sh, progress, ticket, tid, lane, ncols, my_changed, my_evals);
    }

    // ---- teardown: count changes; the last CTA out resets the ticket for the next launch ----------
    unsigned wsum = my_changed, esum = my_evals;
    for (int o = 16; o > 0; o >>= 1) { wsum += __shfl_down_sync(0xffffffffu, wsum, o); esum += __shfl_down_sync(0xffffffffu, esum, o); }
    if (lane == 0 && wsum) atomicAdd(changed, (unsigned long long)wsum);
    if (lane == 0 && esum) atomicAdd(changed + 1, (unsigned long long)esum);
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        unsigned done = atomicAdd(ticket + 1, 1u);
        if (done == gridDim.x - 1) { ticket[0] = 0; ticket[1] = 0; __threadfence(); }
    }
}

// ---- fused launch (the default for the first pass; SDFB_FUSE_PASS=0 turns it off): consecutive sweeps in ONE kernel ---
// Tickets run through the columns of sweep 0 of the launch, then sweep 1, ...; a CTA that has finished its last column
// of one sweep simply takes a column of the next and waits (wait_previous_sweep) until the columns of the previous
// sweep it depends on are complete.  Transitions where only some axes flip overlap; opposite
// directions serialise by themselves.  Every prerequisite holds a lower ticket, so it is running or done: no deadlock.
// Progress words are double-buffered by launch-relative sweep parity.  Cells are read through L2 only (L2OWN).
constexpr int FUSE_MAX = 16;
struct FusedParams {
    int n;                       // sweeps in this launch
    int col_begin[FUSE_MAX + 1]; // first ticket of each sweep
    int flag_stride;             // words between the two progress arrays
    int order_w;                 // ticket order inside a sweep: by w*J + K; 1 = anti-diagonals J + K (see the decode in the kernel)
    unsigned long long *trace;   // SDFB_LINK_TRACE: [2q] = max over columns of ~start time, [2q+1] = max of end time (globaltimer ns)
    ColParams p[FUSE_MAX];
};

static_assert(sizeof(FusedParams) <= 4096, "FusedParams must stay within the classic 4 KB kernel parameter space");

// LINK = exact multi-GPU mode: the slab's neighbours run the same launch on their GPUs; the first K block of every sweep
// waits for (and reads) what the upstream neighbour's last K block hands over, the last K block hands its own boundary
// plane to the downstream neighbour (LinkSweep in sdfb_kernels.cuh).  A column still only ever waits for columns with
// lower tickets on its own GPU or for columns of the SAME sweep on the upstream GPU, and the k direction orders the GPUs
// within a sweep, so the dependency graph over (sweep, position in flow order, ticket) stays acyclic: no deadlock as
// long as every GPU's launch is eventually resident (one launch per GPU, or capped grids when slabs share a GPU).
template <int MINB, bool LINK>
__global__ void __launch_bounds__(NTHREADS, minb_bound(MINB))
k_sweep_columns_fused(uint64_t *__restrict__ cells, const TriRec *__restrict__ rec, const __grid_constant__ FusedParams FP,
                      uint32_t *__restrict__ progress, uint32_t *__restrict__ ticket, unsigned long long *__restrict__ changed)
{
    __shared__ ColShared sh;
    const int tid = threadIdx.x, lane = tid & 31;
    const int ncols = FP.col_begin[FP.n];
    unsigned my_changed = 0, my_evals = 0;
    int q = 0;
    for (;;) {
        if (tid == 0) { sh.col = (int)atomicAdd(ticket, 1u); sh.done = 0; sh.qn = 0; }
        __syncthreads();
        const int tk_all = sh.col;
        if (tk_all >= ncols) break;
        while (tk_all >= FP.col_begin[q + 1]) ++q;                       // tickets only grow
        const ColParams &P = FP.p[q];
        const int tk = tk_all - FP.col_begin[q];
        int J, K;
        if (FP.order_w > 1) {
            // Tickets ordered by the key w*J + K (ties: J ascending).  w = 1 is the anti-diagonal order; a larger w brings
            // the columns of the LAST K block -- the ones a linked slab above waits for -- forward: the neighbour trails
            // this slab by K_last / w rows of J instead of K_last anti-diagonals, at the price of more columns that hold a
            // slot before their producers are far enough (w >= NK is row by row).  Both in-sweep prerequisites, (J-1,K) with
            // key - w and (J,K-1) with key - 1, still hold lower tickets.
            const int w = FP.order_w;
            int g = 0, rem = tk;
            for (;;) {
                int jlo = g - (P.NK - 1) > 0 ? (g - (P.NK - 1) + w - 1) / w : 0, jhi = min(g / w, P.NJ - 1);
                int cnt = jhi - jlo + 1;
                if (cnt > 0) { if (rem < cnt) { J = jlo + rem; K = g - w * J; break; } rem -= cnt; }
                ++g;
            }
        } else {
            int d = 0, rem = tk;
            for (;;) {
                int lo = max(0, d - (P.NK - 1)), hi = min(d, P.NJ - 1);
                int cnt = hi - lo + 1;
                if (rem < cnt) { J = lo + rem; K = d - J; break; }
                rem -= cnt; ++d;
            }
        }
        uint32_t *flags = progress + (q & 1) * FP.flag_stride;
        const int rj0 = 1 + J * EJ, rk0 = P.rk_first + K * EK;
        if (LINK && FP.trace && tid == 0) atomicMax(&FP.trace[2 * q], ~global_timer_ns());
        // evaluation-heavy sweeps (the first pass) share one evaluation queue per column; light ones (the second pass, when
        // it runs in this launch: linked slabs) keep warp-private queues and skip the queue barriers -- uniform per column
        if (tid < NCOMPUTE) {
            if (P.cta_queue) compute_column<true, true, LINK>(cells, rec, P, sh, tid, rj0, rk0, my_changed, my_evals);
            else compute_column<false, true, LINK>(cells, rec, P, sh, tid, rj0, rk0, my_changed, my_evals);
        } else if (tid < NSTEPPERS) {
            if (MERGE) {
                // the halo warp is the sync warp too: previous sweep's prerequisites, release of the compute lanes, then the column
                if (q > 0) wait_previous_sweep(P, FP.p[q - 1], progress + ((q - 1) & 1) * FP.flag_stride, tid - NCOMPUTE, J, K);
                bar_start_arrive();
                const uint32_t *prog_left = (J > 0) ? &flags[K * P.NJ + (J - 1)] : nullptr;
                const uint32_t *prog_down = (K > 0) ? &flags[(K - 1) * P.NJ + J] : nullptr;
                const unsigned long long *link_down = (LINK && K == 0 && P.link.flag_src) ? &P.link.flag_src[J] : nullptr;
                unsigned long long *link_mine = (LINK && K == P.NK - 1 && P.link.flag_dst) ? &P.link.flag_dst[J] : nullptr;
                if (P.cta_queue) halo_column<true, LINK>(cells, rec, P, sh, tid - NCOMPUTE, rj0, rk0, my_evals, prog_left, prog_down, &flags[K * P.NJ + J], link_down, link_mine);
                else halo_column<false, LINK>(cells, rec, P, sh, tid - NCOMPUTE, rj0, rk0, my_evals, prog_left, prog_down, &flags[K * P.NJ + J], link_down, link_mine);
            } else {
                if (P.cta_queue) halo_column<true, LINK>(cells, rec, P, sh, tid - NCOMPUTE, rj0, rk0, my_evals);
                else halo_column<false, LINK>(cells, rec, P, sh, tid - NCOMPUTE, rj0, rk0, my_evals);
            }
        } else if (!MERGE && tid < NSTEPPERS + 32) {
            if (q > 0) wait_previous_sweep(P, FP.p[q - 1], progress + ((q - 1) & 1) * FP.flag_stride, lane, J, K);
            bar_start_arrive();
            const uint32_t *prog_left = (J > 0) ? &flags[K * P.NJ + (J - 1)] : nullptr;
            const uint32_t *prog_down = (K > 0) ? &flags[(K - 1) * P.NJ + J] : nullptr;
            const unsigned long long *link_down = (LINK && K == 0 && P.link.flag_src) ? &P.link.flag_src[J] : nullptr;
            unsigned long long *link_mine = (LINK && K == P.NK - 1 && P.link.flag_dst) ? &P.link.flag_dst[J] : nullptr;
            sync_column<LINK>(P, sh, lane, prog_left, prog_down, &flags[K * P.NJ + J], link_down, link_mine);
        }
        __syncthreads();
        if (LINK && FP.trace && tid == 0) atomicMax(&FP.trace[2 * q + 1], global_timer_ns());
    }
    unsigned wsum = my_changed, esum = my_evals;
    for (int o = 16; o > 0; o >>= 1) { wsum += __shfl_down_sync(0xffffffffu, wsum, o); esum += __shfl_down_sync(0xffffffffu, esum, o); }
    if (lane == 0 && wsum) atomicAdd(changed, (unsigned long long)wsum);
    if (lane == 0 && esum) atomicAdd(changed + 1, (unsigned long long)esum);
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        unsigned done = atomicAdd(ticket + 1, 1u);
        if (done == gridDim.x - 1) { ticket[0] = 0; ticket[1] = 0; __threadfence(); }
    }
}

bool fill_col_params(ColParams &P, const Grid &g, int sweep_index, uint32_t epoch)
{
    P = ColParams{};
    P.g = g;
    P.run_if = nullptr;
    P.sd = SweepDir::of(sweep_index);
    int rk_lo, rk_hi;
    if (!P.sd.owned_rk_range(g, rk_lo, rk_hi)) return false;
    if (g.ni < 2 || g.nj < 2) return false;
    P.rk_first = rk_lo; P.rk_last = rk_hi;
    P.NJ = (g.nj - 1 + EJ - 1) / EJ;
    P.NK = (rk_hi - rk_lo + 1 + EK - 1) / EK;
    P.steps = (g.ni + EJ + EK - 2 + SHIFT + PUBLISH - 1) / PUBLISH * PUBLISH;
    P.stamp = (uint32_t)min(sweep_index + 1, 31);
    P.epoch = epoch;
    memo_last_table(sweep_index, P.sd, P.last);
    return true;
}

}  // namespace

// Sweeps first .. first+count-1 (all with the column-wide evaluation queue) in one launch.
// Returns the number of launches (1), or 0 if this grid cannot be fused (then the caller launches sweep by sweep);
// *epoch is advanced by one per sweep.  link != nullptr: exact multi-GPU mode (the caller has made sure every sweep of
// the range updates at least one plane of this slab and that first + count <= LINK_SWEEPS).
int COLS_FN(launch_sweep_columns_fused)(uint64_t *cells, const TriRec *rec, const Grid &g, int first, int count,
                               unsigned long long *changed, uint32_t *progress, size_t progress_words, uint32_t *epoch,
                               cudaStream_t st, const Tuning &tun, int max_ctas, const LinkState *link)
{
    if (WG || count < (link ? 1 : 2) || count > FUSE_MAX) return 0;
    if (link && first + count > LINK_SWEEPS) return 0;
    FusedParams FP{};
    FP.n = count;
    FP.trace = link && link->trace ? link->trace + 2 * first : nullptr;
    FP.order_w = tun.order_w >= 1 ? tun.order_w : 1;
    FP.flag_stride = (int)((progress_words - 4) / 2);
    FP.col_begin[0] = 0;
    for (int q = 0; q < count; ++q) {
        ColParams &P = FP.p[q];
        if (!fill_col_params(P, g, first + q, *epoch + 1 + (uint32_t)q)) return 0;      // a sweep with nothing to update: do not fuse
        if ((size_t)P.NJ * P.NK > (size_t)FP.flag_stride) return 0;
        FP.col_begin[q + 1] = FP.col_begin[q] + P.NJ * P.NK;
        P.cta_queue = tun.cta_queue >= 0 ? (tun.cta_queue != 0) : (first + q < tun.cta_queue_until);
        if (link) {
            // upstream side of this sweep: the slab below for dk > 0, above for dk < 0 (if there is one); downstream: the other
            const int s = first + q, up = P.sd.dk > 0 ? 0 : 1, down = 1 - up;
            const bool has_up = up == 0 ? g.k_lo > 0 : g.k_hi < g.nk, has_down = down == 0 ? g.k_lo > 0 : g.k_hi < g.nk;
            const size_t plane = (size_t)g.plane();
            if (has_up) { P.link.halo_src = link->in_halo + (size_t)s * plane; P.link.flag_src = link->in_flags + (size_t)s * link->NJ; }
            if (has_down) {
                if (!link->peer_halo[down] || !link->peer_flags[down]) return 0;           // neighbour not linked: the caller reports it
                P.link.halo_dst = link->peer_halo[down] + (size_t)s * plane;
                P.link.flag_dst = link->peer_flags[down] + (size_t)s * link->NJ;
            }
            // timing experiments (SDFB_LINK_DEBUG, results are wrong): local stores / no upstream wait / device-scope fence
            if ((tun.link_debug & 1) && has_down) P.link.halo_dst = link->in_halo + (size_t)((s + 8) % LINK_SWEEPS) * plane;
            if ((tun.link_debug & 4) && has_up) P.link.flag_src = nullptr;
            P.link_gpu_fence = (tun.link_debug & 2) ? 1 : 0;
            P.run_base = link->run << 32;
            P.link_timeout_ns = tun.link_timeout_s > 0 ? (unsigned long long)tun.link_timeout_s * 1000000000ull : 0ull;
        }
    }
    *epoch += (uint32_t)count;
    int dev = 0, sms = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int minb = ((int64_t)g.ni * (g.nj - 1) * (FP.p[0].rk_last - FP.p[0].rk_first + 1) >= ((int64_t)300 << 20)) ? 4 : MINB_SMALL;
    if (tun.minb) minb = tun.minb >= 4 ? 4 : 3;
    using kern_t = void (*)(uint64_t *, const TriRec *, const FusedParams, uint32_t *, uint32_t *, unsigned long long *);
    const kern_t kern = link ? (minb == 4 ? k_sweep_columns_fused<4, true> : k_sweep_columns_fused<3, true>)
                             : (minb == 4 ? k_sweep_columns_fused<4, false> : k_sweep_columns_fused<3, false>);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NTHREADS, 0);
    if (occ < 1) occ = 1;
    if (tun.max_occ > 0 && occ > tun.max_occ) occ = tun.max_occ;
    int grid = sms * occ;
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    // small grids: a sweep has fewer columns than the device has CTA slots, and CTAs beyond that would only sit on
    // tickets of later sweeps and poll; keep a quarter more than one sweep can use so that the next sweep starts at once
    int maxcols = 0;
    for (int q = 0; q < count; ++q) maxcols = max(maxcols, FP.p[q].NJ * FP.p[q].NK);
    if (grid > maxcols + maxcols / 4 + 1) grid = maxcols + maxcols / 4 + 1;
    if (grid > FP.col_begin[count]) grid = FP.col_begin[count];
    kern<<<grid, NTHREADS, 0, st>>>(cells, rec, FP, progress + 4, progress, changed);
    return 1;
}

size_t COLS_FN(link_flag_words_per_sweep)(const Grid &g) { return (size_t)(g.nj - 1 + EJ - 1) / EJ + 1; }

// progress: [2] ticket words + [1] epoch counter slot (host side keeps the epoch) + 2 x NJ*NK flags (the second array is
// used by fused launches only)
size_t COLS_FN(sweep_columns_progress_words)(const Grid &g)
{
    size_t NJ = (size_t)(g.nj + EJ - 1) / EJ + 1, NK = (size_t)(g.nkl() + EK - 1) / EK + 1;
    return 4 + 2 * NJ * NK;
}

int COLS_FN(launch_sweep_columns)(uint64_t *cells, const TriRec *rec, const Grid &g, int sweep_index,
                         unsigned long long *changed, uint32_t *progress, uint32_t epoch, cudaStream_t st,
                         const Tuning &tun, const unsigned int *run_if, int max_ctas)
{
    ColParams P{};
    P.g = g;
    P.run_if = run_if;
    P.sd = SweepDir::of(sweep_index);
    int rk_lo, rk_hi;
    if (!P.sd.owned_rk_range(g, rk_lo, rk_hi)) return 0;
    if (g.ni < 2 || g.nj < 2) return 0;
    P.rk_first = rk_lo; P.rk_last = rk_hi;
    P.NJ = (g.nj - 1 + EJ - 1) / EJ;
    P.NK = (rk_hi - rk_lo + 1 + EK - 1) / EK;
    P.steps = (g.ni + EJ + EK - 2 + SHIFT + PUBLISH - 1) / PUBLISH * PUBLISH;   // whole chunks (PUBLISH is even: the step loops are unrolled by two)
    P.stamp = (uint32_t)min(sweep_index + 1, 31);
    // the epoch grows with every launch on a plan between resets of the progress array (host side)
    P.epoch = epoch;
    memo_last_table(sweep_index, P.sd, P.last);
#ifdef SDFB_TRACE
    static unsigned long long *trace_buf = nullptr;
    const size_t trace_n = (size_t)11 * 8192 * 8;
    if (getenv("SDFB_TRACE") && P.steps <= 8192) {
        if (!trace_buf) cudaMalloc(&trace_buf, trace_n * sizeof(unsigned long long));
        cudaMemsetAsync(trace_buf, 0, trace_n * sizeof(unsigned long long), st);
        P.trace = trace_buf;
        P.trace_col = getenv("SDFB_TRACE_COL") ? atoi(getenv("SDFB_TRACE_COL")) : (P.NJ * P.NK) / 2;
    }
#endif
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // evaluation-heavy sweeps (the first pass) balance the distance evaluations over the whole column
    const bool cta_queue = tun.cta_queue >= 0 ? tun.cta_queue != 0 : (sweep_index < tun.cta_queue_until);
    int occ = 1;
    // register bound by the amount of work per launch (see k_sweep_columns)
    int minb = ((int64_t)g.ni * (g.nj - 1) * (rk_hi - rk_lo + 1) >= ((int64_t)300 << 20)) ? 4 : MINB_SMALL;
    if (tun.minb) minb = tun.minb >= 4 ? 4 : 3;
    using kern_t = void (*)(uint64_t *, const TriRec *, ColParams, uint32_t *, uint32_t *, unsigned long long *);
    const kern_t kern = cta_queue ? (minb == 4 ? k_sweep_columns<true, 4> : k_sweep_columns<true, 3>)
                                  : (minb == 4 ? k_sweep_columns<false, 4> : k_sweep_columns<false, 3>);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NTHREADS, 0);
    if (occ < 1) occ = 1;
    if (tun.max_occ > 0) occ = min(occ, tun.max_occ);   // experiment knob
    int grid = sms * occ;
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;            // batch mode: several plans share the device
    int ncols = P.NJ * P.NK;
    if (grid > ncols) grid = ncols;
    kern<<<grid, NTHREADS, 0, st>>>(cells, rec, P, progress + 4, progress, changed);
#ifdef SDFB_TRACE
    if (P.trace) {
        cudaStreamSynchronize(st);
        unsigned long long *h = (unsigned long long *)malloc(trace_n * sizeof(unsigned long long));
        cudaMemcpy(h, trace_buf, trace_n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        char name[256];
        snprintf(name, sizeof(name), "%s.%d.bin", getenv("SDFB_TRACE"), sweep_index);
        FILE *f = fopen(name, "wb");
        if (f) { fwrite(h, sizeof(unsigned long long), trace_n, f); fclose(f); }
        free(h);
    }
#endif
    return 1;
}

}  // namespace sdfb
