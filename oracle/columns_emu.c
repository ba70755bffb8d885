/*
 * columns_emu.c -- CPU emulation of the pipelined-column sweep schedule (TEST INFRASTRUCTURE ONLY).
 *
 * Mirrors sdfgen_b200/csrc/sdfb_sweep_columns.cu statement by statement (same lane numbering, ring
 * slots, halo-lane prefetch, candidate filter with the stamp memo, owner replay) but runs the
 * columns one after another in ticket order on the CPU, so the index arithmetic, the ring timing and
 * the memo rule of the CUDA kernel can be checked against the serial oracle without a GPU
 * (tests/test_columns_emu.py).  It also ASSERTS the progress-flag arithmetic: every word a halo lane
 * loads must have been produced at a step the consumer's wait condition has already covered.
 *
 * It cannot check memory-model behaviour (fences, cache bypass); the GPU parity tests do that.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

float sdfo_point_triangle_distance(const float *x0, const float *x1, const float *x2, const float *x3);

/* column cross-section: the library builds the schedule twice, 8 x 16 (default here) and 8 x 12 (sdfb_sweep_columns_ek12.cu) */
static int EJ = 8, EK = 16;
void sdfo_emu_set_column_shape(int ej, int ek) { if (ej >= 1 && ek >= 1 && ej * ek <= 256 && (ej * ek) % 32 == 0) { EJ = ej; EK = ek; } }
#define NCOMPUTE (EJ*EK)
#define NLANES (NCOMPUTE + 64)
#define PUBLISH 2
#define RING 2
#define SHIFT 2
#define TRI_MASK 0x07ffffffu
#define TRI_NONE 0x07ffffffu

typedef struct { int ni, nj, nk, k_lo, k_hi; float dx, ox, oy, oz; } egrid;

static const int DIRS[8][3] = { {+1,+1,+1}, {-1,-1,-1}, {+1,+1,-1}, {-1,-1,+1}, {+1,-1,+1}, {-1,+1,-1}, {+1,-1,-1}, {-1,+1,+1} };

static int64_t cidx(const egrid *g, int i, int j, int k) { return (int64_t)i + (int64_t)g->ni*((int64_t)j + (int64_t)g->nj*(int64_t)(k - g->k_lo + 1)); }
static int ring_idx(int slot, int a, int b) { return slot*((EK+1)*(EJ+1)) + (b+1)*(EJ+1) + (a+1); }
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* failure counter readable from Python */
static long emu_flag_violations = 0;
long sdfo_emu_flag_violations(void) { return emu_flag_violations; }

/*
 * cells_phi / cells_lo: (nkl+2) planes each (halo plane first and last), i fastest, updated in place.
 * write_step: scratch of the same size (int32), records the producer step of every cell per sweep.
 * Returns the number of evaluations performed (for DESIGN.md), or -1 on allocation failure.
 */
long sdfo_emu_sweep_columns(const uint32_t *tri, const float *x, float *cells_phi, uint32_t *cells_lo,
                            const float origin[3], float dx, int ni, int nj, int nk, int k_lo, int k_hi,
                            int sweep_index, long *changed_out)
{
    egrid G = { ni, nj, nk, k_lo, k_hi, dx, origin[0], origin[1], origin[2] };
    const egrid *g = &G;
    const int di = DIRS[sweep_index % 8][0], dj = DIRS[sweep_index % 8][1], dk = DIRS[sweep_index % 8][2];
    #define ABS_I(r) (di > 0 ? (r) : ni - 1 - (r))
    #define ABS_J(r) (dj > 0 ? (r) : nj - 1 - (r))
    #define ABS_K(r) (dk > 0 ? (r) : nk - 1 - (r))
    int ra = dk > 0 ? k_lo : nk - 1 - k_lo, rb = dk > 0 ? k_hi - 1 : nk - 1 - (k_hi - 1);
    int rk_first = imin(ra, rb), rk_last = imax(ra, rb);
    if (rk_first < 1) rk_first = 1;
    if (rk_first > rk_last || ni < 2 || nj < 2) { if (changed_out) *changed_out = 0; return 0; }
    const int NJ = (nj - 1 + EJ - 1) / EJ, NK = (rk_last - rk_first + 1 + EK - 1) / EK;
    const int steps = (ni + EJ + EK - 2 + SHIFT + 1) & ~1;
    const uint32_t stamp = (uint32_t)imin(sweep_index + 1, 31);
    uint8_t last[8][7];
    for (int c = 0; c < 8; ++c) for (int m = 0; m < 7; ++m) {
        last[c][m] = 0;
        if (sweep_index + 1 > 31) continue;
        int ci = (m == 0 || m == 2 || m == 4 || m == 6), cj = (m == 1 || m == 2 || m == 5 || m == 6), ck = (m >= 3);
        for (int e = sweep_index - 1; e >= 0; --e) {
            const int *d = DIRS[e % 8];
            int same_i = d[0] == di, same_j = d[1] == dj, same_k = d[2] == dk;
            if ((!(ci || (c & 1)) || same_i) && (!(cj || (c & 2)) || same_j) && (!(ck || (c & 4)) || same_k)) { last[c][m] = (uint8_t)(e + 1); break; }
        }
    }
    const int64_t ncell = (int64_t)ni * nj * (k_hi - k_lo + 2);
    int32_t *write_step = (int32_t*)malloc(sizeof(int32_t) * (size_t)ncell);
    if (!write_step) return -1;
    for (int64_t c = 0; c < ncell; ++c) write_step[c] = -1;     /* -1: not written in this sweep */
    const int64_t si = di;
    long evals = 0, changed = 0;

    uint32_t ring[RING * (EK + 1) * (EJ + 1)];
    static uint32_t q_ent[8][7 * 32];
    static float q_d[8][7 * 32];

    const int ncols = NJ * NK;
    for (int tk = 0; tk < ncols; ++tk) {
        int J, K;
        { int d = 0, rem = tk; for (;;) { int lo = imax(0, d - (NK - 1)), hi = imin(d, NJ - 1), cnt = hi - lo + 1; if (rem < cnt) { J = lo + rem; K = d - J; break; } rem -= cnt; ++d; } }
        const int rj0 = 1 + J * EJ, rk0 = rk_first + K * EK;
        int A[NLANES], B[NLANES], row_ok[NLANES];
        int64_t c_row[NLANES];
        uint64_t own_next_phi_lo[NLANES];   /* packed like the device: phi bits << 32 | lo */
        uint32_t halo_w1[NLANES], halo_next[NLANES], prev_lo[NLANES], r1_old[NLANES], r3_old[NLANES], r5_old[NLANES], r5_old2[NLANES];
        for (int tid = 0; tid < NLANES; ++tid) {
            int a, b;
            if (tid < NCOMPUTE) { a = tid % EJ; b = tid / EJ; }
            else { int h = tid - NCOMPUTE; if (h <= EK) { a = -1; b = h - 1; } else if (h <= EK + EJ) { a = h - EK - 1; b = -1; } else { a = -2; b = -2; } }
            A[tid] = a; B[tid] = b;
            int rj = rj0 + a, rk = rk0 + b;
            row_ok[tid] = (a > -2) && rj <= nj - 1 && rk <= rk_last;
            c_row[tid] = 0;
            if (row_ok[tid]) {
                int j = ABS_J(rj), k = ABS_K(rk);
                c_row[tid] = cidx(g, ABS_I(0), j, k);
            }
            own_next_phi_lo[tid] = 0; halo_next[tid] = TRI_NONE; prev_lo[tid] = TRI_NONE;
            r1_old[tid] = r3_old[tid] = r5_old[tid] = r5_old2[tid] = TRI_NONE;
            int ri0 = 0 - a - b - SHIFT;
            if (tid < NCOMPUTE && row_ok[tid] && ri0 >= 0 && ri0 <= ni - 1) {
                int64_t c = c_row[tid] + si * ri0; uint32_t pb; memcpy(&pb, &cells_phi[c], 4);
                own_next_phi_lo[tid] = ((uint64_t)pb << 32) | cells_lo[c];
            }
        }
        for (int s0 = 0; s0 < steps; s0 += PUBLISH) {
            const int s1 = imin(s0 + PUBLISH, steps);
            /* what the wait condition of this chunk guarantees about the producers */
            const int need_left = imin(steps, s1 - 1 + EJ + 3), need_down = imin(steps, s1 - 1 + EK + 3);
            if (s0 == 0) for (int tid = NCOMPUTE; tid < NLANES; ++tid) if (row_ok[tid]) {
                int ri0 = 0 - A[tid] - B[tid] - SHIFT;
                halo_next[tid] = (ri0 >= 0 && ri0 <= ni - 1) ? cells_lo[c_row[tid] + si * ri0] : TRI_NONE;
                halo_w1[tid] = (ri0 + 1 >= 0 && ri0 + 1 <= ni - 1) ? cells_lo[c_row[tid] + si * (ri0 + 1)] : TRI_NONE;
            }
            for (int s = s0; s < s1; ++s) {
                const int slot = s & 1, pslot = slot ^ 1;
                /* halo lanes */
                for (int tid = NCOMPUTE; tid < NLANES; ++tid) if (row_ok[tid]) {
                    int a = A[tid], b = B[tid], ri = s - a - b - SHIFT;
                    ring[ring_idx(slot, a, b)] = halo_next[tid];
                    halo_next[tid] = halo_w1[tid];
                    int rin = ri + 2;                       /* the load issued at step s is for virtual step s+2 */
                    if (rin >= 0 && rin <= ni - 1 && s + 2 < steps) {
                        int64_t c = c_row[tid] + si * rin;
                        halo_w1[tid] = cells_lo[c];
                        /* flag arithmetic: the producer step of this voxel must be covered by the wait */
                        int rj = rj0 + a, rk = rk0 + b;
                        if (rin >= 1 && rj >= 1 && rk >= rk_first) {           /* a voxel some column updates */
                            int pa = (rj - 1) % EJ, pb = (rk - rk_first) % EK;
                            int pstep = rin + pa + pb + SHIFT;                          /* step inside its producer column */
                            int guaranteed;
                            if (a == -1 && b >= 0) guaranteed = need_left;
                            else if (b == -1 && a >= 0) guaranteed = need_down;
            /* diagonal column: covered transitively -- the left column ran step need_left-1 only
                               after ITS wait saw the diagonal column at >= need_left-1+EK+2 steps */
                            else guaranteed = imin(steps, need_left - 1 + EK + 3);
                            if (!(pstep < guaranteed)) ++emu_flag_violations;
                        }
                    } else halo_w1[tid] = TRI_NONE;
                }
                /* compute lanes: phase 1, candidates */
                int ncand[NCOMPUTE], upd[NCOMPUTE], in_row_v[NCOMPUTE];
                uint32_t cand[NCOMPUTE][7], cur_v[NCOMPUTE];
                float phi_v[NCOMPUTE];
                for (int tid = 0; tid < NCOMPUTE; ++tid) {
                    int a = A[tid], b = B[tid], ri = s - a - b - SHIFT;
                    int in_row = row_ok[tid] && ri >= 0 && ri <= ni - 1;
                    uint64_t self = own_next_phi_lo[tid];
                    { int rin = ri + 1; if (row_ok[tid] && rin >= 0 && rin <= ni - 1) { int64_t c = c_row[tid] + si * rin; uint32_t pb; memcpy(&pb, &cells_phi[c], 4); own_next_phi_lo[tid] = ((uint64_t)pb << 32) | cells_lo[c]; } }
                    uint32_t cur = (uint32_t)self; uint32_t pb = (uint32_t)(self >> 32); float phi; memcpy(&phi, &pb, 4);
                    uint32_t r1 = TRI_NONE, r3 = TRI_NONE, r5 = TRI_NONE;
                    if (row_ok[tid] && ri >= -1 && ri <= ni - 1) {
                        r1 = ring[ring_idx(pslot, a - 1, b)];
                        r3 = ring[ring_idx(pslot, a, b - 1)];
                        r5 = ring[ring_idx(pslot, a - 1, b - 1)];
                    }
                    int n = 0; int update = in_row && ri >= 1;
                    if (update) {
                        uint32_t nb[7] = { prev_lo[tid], r1, r1_old[tid], r3, r3_old[tid], r5_old[tid], r5_old2[tid] };
                        int i = ABS_I(ri);
                        int i_interior = (i >= 1 && i <= ni - 2);
                        (void)i_interior;
                        int cls = (ri == ni - 1 ? 1 : 0) | (rj0 + a == nj - 1 ? 2 : 0) | (rk0 + b == nk - 1 ? 4 : 0);
                        uint32_t live = 0;
                        for (int m = 0; m < 7; ++m) {
                            uint32_t xw = nb[m];
                            uint32_t thr = last[cls][m] ? ((uint32_t)last[cls][m] + 1u) << 27 : 0u;
                            int keep = ((xw & TRI_MASK) != TRI_NONE) && (((xw ^ cur) & TRI_MASK) != 0) && (xw >= thr);
                            if (keep) live |= 1u << m;
                        }
                        if (live) {
                            for (int m = 1; m < 7; ++m) {
                                int dup = 0;
                                for (int u = 0; u < m; ++u) dup = dup || (((nb[u] ^ nb[m]) & TRI_MASK) == 0);
                                if (dup) live &= ~(1u << m);
                            }
                        }
                        for (int m = 0; m < 7; ++m) { int keep = (live >> m) & 1u; cand[tid][m] = keep ? (nb[m] & TRI_MASK) : TRI_NONE; n += keep; }
                    }
                    r1_old[tid] = r1; r3_old[tid] = r3; r5_old2[tid] = r5_old[tid]; r5_old[tid] = r5;
                    ncand[tid] = n; upd[tid] = update; in_row_v[tid] = in_row; cur_v[tid] = cur; phi_v[tid] = phi;
                }
                /* phase 2..4 per warp */
                for (int w = 0; w < NCOMPUTE / 32; ++w) {
                    int off[32], total = 0;
                    for (int l = 0; l < 32; ++l) { off[l] = total; total += ncand[w * 32 + l]; }
                    if (total > 0) {
                        for (int l = 0; l < 32; ++l) { int tid = w * 32 + l; if (upd[tid]) { int q = off[l]; for (int m = 0; m < 7; ++m) if (cand[tid][m] != TRI_NONE) q_ent[w][q++] = ((uint32_t)l << 27) | cand[tid][m]; } }
                        for (int q = 0; q < total; ++q) {
                            uint32_t e = q_ent[w][q];
                            int ol = (int)(e >> 27), otid = w * 32 + ol, oa = otid % EJ, ob = otid / EJ, ori = s - oa - ob - SHIFT;
                            int oi = ABS_I(ori), oj = ABS_J(rj0 + oa), ok = ABS_K(rk0 + ob);
                            float gx[3] = { oi * dx + origin[0], oj * dx + origin[1], ok * dx + origin[2] };
                            uint32_t t = e & TRI_MASK;
                            q_d[w][q] = sdfo_point_triangle_distance(gx, x + 3 * (size_t)tri[3 * (size_t)t], x + 3 * (size_t)tri[3 * (size_t)t + 1], x + 3 * (size_t)tri[3 * (size_t)t + 2]);
                            ++evals;
                        }
                        for (int l = 0; l < 32; ++l) {
                            int tid = w * 32 + l;
                            if (ncand[tid] > 0) {
                                uint32_t best = TRI_NONE; float phi = phi_v[tid];
                                for (int q = off[l]; q < off[l] + ncand[tid]; ++q) { float d = q_d[w][q]; if (d < phi) { phi = d; best = q_ent[w][q] & TRI_MASK; } }
                                if (best != TRI_NONE) {
                                    int ri = s - A[tid] - B[tid] - SHIFT;
                                    int64_t c = c_row[tid] + si * ri;
                                    cur_v[tid] = (stamp << 27) | best;
                                    cells_phi[c] = phi; cells_lo[c] = cur_v[tid]; write_step[c] = s;
                                    ++changed;
                                }
                            }
                        }
                    }
                }
                for (int tid = 0; tid < NCOMPUTE; ++tid) if (in_row_v[tid]) { ring[ring_idx(slot, A[tid], B[tid])] = cur_v[tid]; prev_lo[tid] = cur_v[tid]; }
            }
        }
    }
    free(write_step);
    if (changed_out) *changed_out = changed;
    return evals;
}
