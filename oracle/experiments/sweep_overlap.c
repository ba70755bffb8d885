/*
 * sweep_overlap.c -- CPU experiment (TEST INFRASTRUCTURE / design study, never linked into the product).
 *
 * Question for the next round: may column (J',K') of sweep s+1 start before sweep s has finished everywhere?
 * Rule under test: a column of sweep s+1 covers the rows (j,k) of a rectangle R, all i.  It writes the cells of R and
 * reads R grown by one row on its upstream sides.  Sweep s must (a) have produced its final values there and (b) not
 * need the pre-(s+1) values of anything sweep s+1 writes; sweep s reads one row around the rows it updates.  Both hold
 * once sweep s has COMPLETED every row of R grown by one in all four directions, i.e. every sweep-s column that
 * intersects the grown rectangle.
 *
 * The program processes whole columns (EJ x EK rows, all i, serial order inside) in this adversarial order: sweep-s
 * columns by ticket (anti-diagonals J+K); after each one, every sweep-(s+1) column that has become ready (its own
 * left/down/diagonal columns of sweep s+1 done, the rule above satisfied) runs at once.  The result must equal the two
 * serial sweeps bit for bit.  It also reports how early the columns of sweep s+1 became ready.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

float sdfo_point_triangle_distance(const float *x0, const float *x1, const float *x2, const float *x3);

static const int DIRS[8][3] = { {+1,+1,+1}, {-1,-1,-1}, {+1,+1,-1}, {-1,-1,+1}, {+1,-1,+1}, {-1,+1,-1}, {+1,-1,-1}, {-1,+1,+1} };

typedef struct { int ni, nj, nk; float dx, o[3]; const uint32_t *tri; const float *x; float *phi; int32_t *ctri; int k_lo, k_hi; } G;

/* sweep parameters as the CUDA launch computes them (sdfb_sweep_common.cuh: owned_rk_range; sdfb_sweep_columns.cu: fill_col_params) */
typedef struct { int di, dj, dk, rk_first, rk_last, NJ, NK, ok; } SP;
static SP sweep_params(const G *g, int s, int EJ, int EK)
{
    SP p; p.di = DIRS[s % 8][0]; p.dj = DIRS[s % 8][1]; p.dk = DIRS[s % 8][2];
    int a = p.dk > 0 ? g->k_lo : g->nk - 1 - g->k_lo, b = p.dk > 0 ? g->k_hi - 1 : g->nk - 1 - (g->k_hi - 1);
    int lo = a < b ? a : b, hi = a < b ? b : a;
    if (lo < 1) lo = 1;
    p.rk_first = lo; p.rk_last = hi; p.ok = lo <= hi && g->ni >= 2 && g->nj >= 2;
    p.NJ = (g->nj - 1 + EJ - 1) / EJ; p.NK = p.ok ? (hi - lo + 1 + EK - 1) / EK : 0;
    return p;
}

/* LITERAL PORT of the device arithmetic of wait_previous_sweep (sdfb_sweep_columns.cu): the columns [Ja,Jb] x [Ka,Kb] of
 * the previous sweep Q that column (J,K) of sweep P waits for; returns 0 if there are none. */
static int prereq_device(const G *g, const SP *P, const SP *Q, int EJ, int EK, int J, int K, int *Ja, int *Jb, int *Ka, int *Kb)
{
    #define MIN(a, b) ((a) < (b) ? (a) : (b))
    #define MAX(a, b) ((a) > (b) ? (a) : (b))
    int rja = 1 + J * EJ, rjb = MIN(rja + EJ - 1, g->nj - 1);
    int rka = P->rk_first + K * EK, rkb = MIN(rka + EK - 1, P->rk_last);
    int j0 = P->dj > 0 ? rja : g->nj - 1 - rjb, j1 = P->dj > 0 ? rjb : g->nj - 1 - rja;
    int k0 = P->dk > 0 ? rka : g->nk - 1 - rkb, k1 = P->dk > 0 ? rkb : g->nk - 1 - rka;
    j0 = MAX(j0 - 1, 0); j1 = MIN(j1 + 1, g->nj - 1); k0 = MAX(k0 - 1, 0); k1 = MIN(k1 + 1, g->nk - 1);
    int qja = Q->dj > 0 ? j0 : g->nj - 1 - j1, qjb = Q->dj > 0 ? j1 : g->nj - 1 - j0;
    int qka = Q->dk > 0 ? k0 : g->nk - 1 - k1, qkb = Q->dk > 0 ? k1 : g->nk - 1 - k0;
    qja = MAX(qja, 1); qjb = MIN(qjb, g->nj - 1); qka = MAX(qka, Q->rk_first); qkb = MIN(qkb, Q->rk_last);
    if (qja > qjb || qka > qkb) return 0;
    *Ja = (qja - 1) / EJ; *Jb = (qjb - 1) / EJ; *Ka = (qka - Q->rk_first) / EK; *Kb = (qkb - Q->rk_first) / EK;
    return 1;
}

static void relax(const G *g, int i, int j, int k, int di, int dj, int dk)      /* cpu_lib/makelevelset3.cpp:104-127 body */
{
    const int64_t c0 = (int64_t)i + (int64_t)g->ni * (j + (int64_t)g->nj * k);
    float gx[3] = { i * g->dx + g->o[0], j * g->dx + g->o[1], k * g->dx + g->o[2] };
    static const int OFF[7][3] = { {1,0,0}, {0,1,0}, {1,1,0}, {0,0,1}, {1,0,1}, {0,1,1}, {1,1,1} };
    for (int m = 0; m < 7; ++m) {
        int64_t c1 = (int64_t)(i - di * OFF[m][0]) + (int64_t)g->ni * ((j - dj * OFF[m][1]) + (int64_t)g->nj * (k - dk * OFF[m][2]));
        int32_t t = g->ctri[c1];
        if (t >= 0) {
            const uint32_t *tv = g->tri + 3 * (size_t)t;
            float d = sdfo_point_triangle_distance(gx, g->x + 3 * (size_t)tv[0], g->x + 3 * (size_t)tv[1], g->x + 3 * (size_t)tv[2]);
            if (d < g->phi[c0]) { g->phi[c0] = d; g->ctri[c0] = t; }
        }
    }
}

static void serial_sweep(const G *g, int s)
{
    int di = DIRS[s % 8][0], dj = DIRS[s % 8][1], dk = DIRS[s % 8][2];
    for (int rk = 1; rk < g->nk; ++rk) for (int rj = 1; rj < g->nj; ++rj) for (int ri = 1; ri < g->ni; ++ri)
        relax(g, di > 0 ? ri : g->ni - 1 - ri, dj > 0 ? rj : g->nj - 1 - rj, dk > 0 ? rk : g->nk - 1 - rk, di, dj, dk);
}

/* absolute row rectangle [j0,j1] x [k0,k1] of column (J,K) of sweep s */
static void col_rect(const G *g, int s, int EJ, int EK, int J, int K, int *j0, int *j1, int *k0, int *k1)
{
    int dj = DIRS[s % 8][1], dk = DIRS[s % 8][2];
    int a = 1 + J * EJ, b = a + EJ - 1; if (b > g->nj - 1) b = g->nj - 1;
    int c = 1 + K * EK, d = c + EK - 1; if (d > g->nk - 1) d = g->nk - 1;
    if (dj > 0) { *j0 = a; *j1 = b; } else { *j0 = g->nj - 1 - b; *j1 = g->nj - 1 - a; }
    if (dk > 0) { *k0 = c; *k1 = d; } else { *k0 = g->nk - 1 - d; *k1 = g->nk - 1 - c; }
}

static void run_column(const G *g, int s, int EJ, int EK, int J, int K)
{
    int di = DIRS[s % 8][0], dj = DIRS[s % 8][1], dk = DIRS[s % 8][2];
    for (int rk = 1 + K * EK; rk < 1 + (K + 1) * EK && rk < g->nk; ++rk)
        for (int rj = 1 + J * EJ; rj < 1 + (J + 1) * EJ && rj < g->nj; ++rj)
            for (int ri = 1; ri < g->ni; ++ri)
                relax(g, di > 0 ? ri : g->ni - 1 - ri, dj > 0 ? rj : g->nj - 1 - rj, dk > 0 ? rk : g->nk - 1 - rk, di, dj, dk);
}

/*
 * Runs sweeps s and s+1 on (phi, ctri) in place with the overlapped column order; ref_phi/ref_tri receive the two
 * serial sweeps applied to a copy.  ready_at[c'] = number of sweep-s columns that had completed when column c' of sweep
 * s+1 ran (== NJ*NK means "only after sweep s had finished").  Returns 0 if the overlapped result equals the serial one.
 */
int sweep_overlap_run(const uint32_t *tri, const float *x, float *phi, int32_t *ctri, float *ref_phi, int32_t *ref_tri,
                      const float origin[3], float dx, int ni, int nj, int nk, int EJ, int EK, int s, int32_t *ready_at)
{
    const int64_t V = (int64_t)ni * nj * nk;
    G g = { ni, nj, nk, dx, { origin[0], origin[1], origin[2] }, tri, x, ref_phi, ref_tri };
    memcpy(ref_phi, phi, sizeof(float) * V); memcpy(ref_tri, ctri, 4 * V);
    serial_sweep(&g, s); serial_sweep(&g, s + 1);
    g.phi = phi; g.ctri = ctri;
    const int NJ = (nj - 1 + EJ - 1) / EJ, NK = (nk - 1 + EK - 1) / EK, NC = NJ * NK;
    uint8_t *done0 = calloc(NC, 1), *done1 = calloc(NC, 1);
    /* prerequisites of every sweep-(s+1) column in terms of sweep-s columns: [Ja,Jb] x [Ka,Kb] */
    int *pre = malloc(sizeof(int) * 4 * NC);
    for (int K = 0; K < NK; ++K) for (int J = 0; J < NJ; ++J) {
        int j0, j1, k0, k1; col_rect(&g, s + 1, EJ, EK, J, K, &j0, &j1, &k0, &k1);
        j0 -= 1; j1 += 1; k0 -= 1; k1 += 1;                                  /* grown by one row in all four directions */
        if (j0 < 0) j0 = 0; if (j1 > nj - 1) j1 = nj - 1; if (k0 < 0) k0 = 0; if (k1 > nk - 1) k1 = nk - 1;
        /* rows -> sweep-s relative rows -> sweep-s columns (relative row 0 is never updated: no column) */
        int dj = DIRS[s % 8][1], dk = DIRS[s % 8][2];
        int ra = dj > 0 ? j0 : nj - 1 - j1, rb = dj > 0 ? j1 : nj - 1 - j0;
        int rc = dk > 0 ? k0 : nk - 1 - k1, rd = dk > 0 ? k1 : nk - 1 - k0;
        if (ra < 1) ra = 1; if (rc < 1) rc = 1;
        int *p = pre + 4 * (K * NJ + J);
        p[0] = (ra - 1) / EJ; p[1] = rb >= 1 ? (rb - 1) / EJ : -1; p[2] = (rc - 1) / EK; p[3] = rd >= 1 ? (rd - 1) / EK : -1;
    }
    int completed0 = 0;
    for (int d = 0; d <= NJ + NK - 2; ++d) for (int J = 0; J < NJ; ++J) {
        int K = d - J; if (K < 0 || K >= NK) continue;
        run_column(&g, s, EJ, EK, J, K); done0[K * NJ + J] = 1; ++completed0;
        /* run every sweep-(s+1) column that is ready now; repeat until none is (anti-diagonal order keeps it cheap) */
        for (int again = 1; again; ) {
            again = 0;
            for (int K1 = 0; K1 < NK; ++K1) for (int J1 = 0; J1 < NJ; ++J1) {
                int c1 = K1 * NJ + J1;
                if (done1[c1]) continue;
                if (J1 > 0 && !done1[c1 - 1]) continue;
                if (K1 > 0 && !done1[c1 - NJ]) continue;
                if (J1 > 0 && K1 > 0 && !done1[c1 - NJ - 1]) continue;
                const int *p = pre + 4 * c1; int ok = 1;
                for (int Kq = p[2]; Kq <= p[3] && ok; ++Kq) for (int Jq = p[0]; Jq <= p[1]; ++Jq) if (!done0[Kq * NJ + Jq]) { ok = 0; break; }
                if (!ok) continue;
                run_column(&g, s + 1, EJ, EK, J1, K1); done1[c1] = 1; ready_at[c1] = completed0; again = 1;
            }
        }
    }
    int missing = 0;
    for (int c1 = 0; c1 < NC; ++c1) missing += !done1[c1];
    free(done0); free(done1); free(pre);
    if (missing) return -1;
    return (memcmp(phi, ref_phi, sizeof(float) * V) || memcmp(ctri, ref_tri, 4 * V)) ? 1 : 0;
}

/* ---- emulation of the fused launch (k_sweep_columns_fused): sweeps first .. first+count-1 on the k-slab [k_lo,k_hi) ----
 * Tickets run through the columns of the sweeps in launch order; W "CTAs" hold the W lowest untaken tickets; among them
 * a READY column is picked (the highest ticket first: the most eager overlap) and run atomically.  Ready = in-sweep
 * left / down / diagonal columns complete + the prerequisites of prereq_device() complete in the PREVIOUS sweep's
 * progress array; the arrays are double-buffered by launch-relative parity exactly like the device's.  Checks: no
 * deadlock, the buffer-reuse invariant (when a column of sweep q+2 runs, sweep q is complete everywhere), and the result
 * against the serial sweeps restricted to the slab (halo planes frozen).  Returns 0 ok, 1 differs, -1 deadlock,
 * -2 invariant violated, -3 a sweep has nothing to update (the device declines to fuse). */
static void run_column_sp(const G *g, const SP *p, int EJ, int EK, int J, int K)
{
    for (int rk = p->rk_first + K * EK; rk < p->rk_first + (K + 1) * EK && rk <= p->rk_last; ++rk)
        for (int rj = 1 + J * EJ; rj < 1 + (J + 1) * EJ && rj < g->nj; ++rj)
            for (int ri = 1; ri < g->ni; ++ri)
                relax(g, p->di > 0 ? ri : g->ni - 1 - ri, p->dj > 0 ? rj : g->nj - 1 - rj, p->dk > 0 ? rk : g->nk - 1 - rk, p->di, p->dj, p->dk);
}

int fused_emulation_run(const uint32_t *tri, const float *x, float *phi, int32_t *ctri, float *ref_phi, int32_t *ref_tri,
                        const float origin[3], float dx, int ni, int nj, int nk, int k_lo, int k_hi, int EJ, int EK,
                        int first, int count, int W, int64_t *early_out)
{
    const int64_t V = (int64_t)ni * nj * nk;
    G g = { ni, nj, nk, dx, { origin[0], origin[1], origin[2] }, tri, x, ref_phi, ref_tri, k_lo, k_hi };
    SP sp[8]; int begin[9]; begin[0] = 0;
    if (count < 2 || count > 8) return -3;
    for (int q = 0; q < count; ++q) { sp[q] = sweep_params(&g, first + q, EJ, EK); if (!sp[q].ok) return -3; begin[q + 1] = begin[q] + sp[q].NJ * sp[q].NK; }
    /* serial reference on the slab */
    memcpy(ref_phi, phi, sizeof(float) * V); memcpy(ref_tri, ctri, 4 * V);
    for (int q = 0; q < count; ++q)
        for (int rk = sp[q].rk_first; rk <= sp[q].rk_last; ++rk) for (int rj = 1; rj < nj; ++rj) for (int ri = 1; ri < ni; ++ri)
            relax(&g, sp[q].di > 0 ? ri : ni - 1 - ri, sp[q].dj > 0 ? rj : nj - 1 - rj, sp[q].dk > 0 ? rk : nk - 1 - rk, sp[q].di, sp[q].dj, sp[q].dk);
    g.phi = phi; g.ctri = ctri;
    int stride = 0;
    for (int q = 0; q < count; ++q) if (sp[q].NJ * sp[q].NK > stride) stride = sp[q].NJ * sp[q].NK;
    uint32_t *flags = calloc((size_t)2 * stride, 4);            /* epoch*2+1 = complete */
    int *completed = calloc(count, sizeof(int));
    int total = begin[count], next_ticket = 0, ntaken = 0, *taken = malloc(sizeof(int) * W), ndone = 0, rc = 0;
    int64_t early = 0;
    while (ndone < total && rc == 0) {
        while (ntaken < W && next_ticket < total) taken[ntaken++] = next_ticket++;
        int pick = -1;
        for (int t = 0; t < ntaken; ++t) {
            int tk = taken[t], q = 0; while (tk >= begin[q + 1]) ++q;
            const SP *P = &sp[q];
            int loc = tk - begin[q], J, K, d = 0, rem = loc;
            for (;;) { int lo = d - (P->NK - 1) > 0 ? d - (P->NK - 1) : 0, hi = d < P->NJ - 1 ? d : P->NJ - 1, cnt = hi - lo + 1; if (rem < cnt) { J = lo + rem; K = d - J; break; } rem -= cnt; ++d; }
            const uint32_t *fl = flags + (q & 1) * stride, mine = (uint32_t)(first + q + 1) * 2 + 1;
            int ok = 1;
            if (J > 0 && fl[K * P->NJ + J - 1] < mine) ok = 0;
            if (K > 0 && fl[(K - 1) * P->NJ + J] < mine) ok = 0;
            if (ok && q > 0) {
                const SP *Q = &sp[q - 1];
                const uint32_t *pf = flags + ((q - 1) & 1) * stride, need = (uint32_t)(first + q) * 2 + 1;
                int Ja, Jb, Ka, Kb;
                if (prereq_device(&g, P, Q, EJ, EK, J, K, &Ja, &Jb, &Ka, &Kb))
                    for (int Kq = Ka; Kq <= Kb && ok; ++Kq) for (int Jq = Ja; Jq <= Jb; ++Jq) if (pf[Kq * Q->NJ + Jq] < need) { ok = 0; break; }
            }
            if (ok && (pick < 0 || taken[t] > taken[pick])) pick = t;
        }
        if (pick < 0) { rc = -1; break; }
        int tk = taken[pick], q = 0; while (tk >= begin[q + 1]) ++q;
        const SP *P = &sp[q];
        int loc = tk - begin[q], J, K, d = 0, rem = loc;
        for (;;) { int lo = d - (P->NK - 1) > 0 ? d - (P->NK - 1) : 0, hi = d < P->NJ - 1 ? d : P->NJ - 1, cnt = hi - lo + 1; if (rem < cnt) { J = lo + rem; K = d - J; break; } rem -= cnt; ++d; }
        if (q >= 2 && completed[q - 2] != sp[q - 2].NJ * sp[q - 2].NK) { rc = -2; break; }       /* buffer-reuse invariant */
        if (q >= 1 && completed[q - 1] != sp[q - 1].NJ * sp[q - 1].NK) ++early;
        run_column_sp(&g, P, EJ, EK, J, K);
        flags[(q & 1) * stride + K * P->NJ + J] = (uint32_t)(first + q + 1) * 2 + 1;
        ++completed[q]; ++ndone;
        taken[pick] = taken[--ntaken];
    }
    if (early_out) *early_out = early;
    free(flags); free(completed); free(taken);
    if (rc) return rc;
    return (memcmp(phi, ref_phi, sizeof(float) * V) || memcmp(ctri, ref_tri, 4 * V)) ? 1 : 0;
}

/* ---- emulation of the LINKED multi-GPU launch (k_sweep_columns_fused<.., LINK = true> on every slab) ---------------
 * nslabs k-slabs [kb[r], kb[r+1]) share the global cell arrays (each writes only its own planes) but read the plane
 * below their slab (in sweep direction) ONLY from their per-sweep inbound buffer, as the device does (LinkSweep in
 * sdfgen_b200/csrc/sdfb_kernels.cuh): a column of the last K block pushes its rows of the slab's boundary plane (plus row
 * rj = 0 when J = 0) to the downstream slab's buffer of that sweep and then raises the per-column flag there; a column
 * of the first K block waits for the flag of the column below it.  Every slab runs its own fused ticket sequence over
 * sweeps 0..count-1 with W "CTA" slots and its own double-buffered progress words; a seeded random choice among ALL
 * ready columns of ALL slabs drives the interleaving.  Checked: no deadlock, never a read of a cell that was not
 * pushed in this sweep (buffers start poisoned), bit equality with the serial sweeps of the whole grid.
 * Returns 0 ok, 1 differs, -1 deadlock, -2 buffer-reuse invariant, -3 a slab declines a sweep, -4 poisoned read. */
#define POISON_TRI (-77)
typedef struct {
    G g; SP sp[16]; int begin[17]; uint32_t *flags; int stride; int *completed; int next_ticket, ntaken, *taken, ndone, total;
    float *in_phi; int32_t *in_tri; uint8_t *in_flag;          /* [16][plane], [16][plane], [16][NJ] */
} Slab;
static int g_poison_reads;

static void relax_linked(const Slab *S, int s, int i, int j, int k, int di, int dj, int dk)
{
    const G *g = &S->g;
    const int64_t plane = (int64_t)g->ni * g->nj;
    const int64_t c0 = (int64_t)i + (int64_t)g->ni * (j + (int64_t)g->nj * k);
    float gx[3] = { i * g->dx + g->o[0], j * g->dx + g->o[1], k * g->dx + g->o[2] };
    static const int OFF[7][3] = { {1,0,0}, {0,1,0}, {1,1,0}, {0,0,1}, {1,0,1}, {0,1,1}, {1,1,1} };
    for (int m = 0; m < 7; ++m) {
        int i1 = i - di * OFF[m][0], j1 = j - dj * OFF[m][1], k1 = k - dk * OFF[m][2];
        int32_t t;
        if (k1 < g->k_lo || k1 >= g->k_hi) {                  /* the plane below the slab: this sweep's inbound buffer */
            t = S->in_tri[(int64_t)s * plane + i1 + (int64_t)g->ni * j1];
            if (t == POISON_TRI) { ++g_poison_reads; continue; }
        } else t = g->ctri[(int64_t)i1 + (int64_t)g->ni * (j1 + (int64_t)g->nj * k1)];
        if (t >= 0) {
            const uint32_t *tv = g->tri + 3 * (size_t)t;
            float d = sdfo_point_triangle_distance(gx, g->x + 3 * (size_t)tv[0], g->x + 3 * (size_t)tv[1], g->x + 3 * (size_t)tv[2]);
            if (d < g->phi[c0]) { g->phi[c0] = d; g->ctri[c0] = t; }
        }
    }
}

static int g_order_w = 1;                   /* tickets ordered by w*J + K: 1 = anti-diagonals, >= NK = row by row (LITERAL PORT of the device decode) */
void linked_emulation_set_order_w(int v) { g_order_w = v < 1 ? 1 : v; }
static void ticket_to_JK(const SP *P, int loc, int *J, int *K)
{
    if (g_order_w > 1) {
        const int w = g_order_w;
        int g = 0, rem = loc;
        for (;;) {
            int jlo = g - (P->NK - 1) > 0 ? (g - (P->NK - 1) + w - 1) / w : 0, jhi = g / w < P->NJ - 1 ? g / w : P->NJ - 1;
            int cnt = jhi - jlo + 1;
            if (cnt > 0) { if (rem < cnt) { *J = jlo + rem; *K = g - w * *J; return; } rem -= cnt; }
            ++g;
        }
    }
    int d = 0, rem = loc;
    for (;;) { int lo = d - (P->NK - 1) > 0 ? d - (P->NK - 1) : 0, hi = d < P->NJ - 1 ? d : P->NJ - 1, cnt = hi - lo + 1; if (rem < cnt) { *J = lo + rem; *K = d - *J; return; } rem -= cnt; ++d; }
}

int linked_emulation_run(const uint32_t *tri, const float *x, float *phi, int32_t *ctri, float *ref_phi, int32_t *ref_tri,
                         const float origin[3], float dx, int ni, int nj, int nk, int nslabs, const int32_t *kb,
                         int EJ, int EK, int count, int W, uint32_t seed, int64_t *overlap_out)
{
    const int64_t V = (int64_t)ni * nj * nk, plane = (int64_t)ni * nj;
    if (count < 1 || count > 16) return -3;
    G gr = { ni, nj, nk, dx, { origin[0], origin[1], origin[2] }, tri, x, ref_phi, ref_tri, 0, nk };
    memcpy(ref_phi, phi, sizeof(float) * V); memcpy(ref_tri, ctri, 4 * V);
    for (int s = 0; s < count; ++s) serial_sweep(&gr, s);
    Slab *S = calloc(nslabs, sizeof(Slab));
    int rc = 0, all_total = 0, all_done = 0;
    for (int r = 0; r < nslabs; ++r) {
        Slab *a = &S[r];
        G g = { ni, nj, nk, dx, { origin[0], origin[1], origin[2] }, tri, x, phi, ctri, kb[r], kb[r + 1] };
        a->g = g; a->begin[0] = 0;
        for (int q = 0; q < count; ++q) { a->sp[q] = sweep_params(&a->g, q, EJ, EK); if (!a->sp[q].ok) rc = -3; a->begin[q + 1] = a->begin[q] + a->sp[q].NJ * a->sp[q].NK; if (a->sp[q].NJ * a->sp[q].NK > a->stride) a->stride = a->sp[q].NJ * a->sp[q].NK; }
        a->flags = calloc((size_t)2 * a->stride + 1, 4); a->completed = calloc(count, sizeof(int)); a->taken = malloc(sizeof(int) * W);
        a->total = a->begin[count]; all_total += a->total;
        a->in_phi = malloc(sizeof(float) * 16 * plane); a->in_tri = malloc(4 * 16 * plane); a->in_flag = calloc((size_t)16 * (a->sp[0].NJ + 1), 1);
        for (int64_t c = 0; c < 16 * plane; ++c) { a->in_phi[c] = -1.f; a->in_tri[c] = POISON_TRI; }
    }
    g_poison_reads = 0;
    int64_t overlap = 0;
    uint32_t rng = seed * 2654435761u + 12345u;
    while (rc == 0 && all_done < all_total) {
        /* candidates: (slab, slot) pairs whose column is ready */
        int nready = 0, pick_r = -1, pick_t = -1;
        for (int r = 0; r < nslabs; ++r) {
            Slab *a = &S[r];
            while (a->ntaken < W && a->next_ticket < a->total) a->taken[a->ntaken++] = a->next_ticket++;
            for (int t = 0; t < a->ntaken; ++t) {
                int tk = a->taken[t], q = 0; while (tk >= a->begin[q + 1]) ++q;
                const SP *P = &a->sp[q]; int J, K; ticket_to_JK(P, tk - a->begin[q], &J, &K);
                const uint32_t *fl = a->flags + (q & 1) * a->stride, mine = (uint32_t)(q + 1) * 2 + 1;
                int ok = 1;
                if (J > 0 && fl[K * P->NJ + J - 1] < mine) ok = 0;
                if (K > 0 && fl[(K - 1) * P->NJ + J] < mine) ok = 0;
                if (ok && q > 0) {
                    const SP *Q = &a->sp[q - 1];
                    const uint32_t *pf = a->flags + ((q - 1) & 1) * a->stride, need = (uint32_t)q * 2 + 1;
                    int Ja, Jb, Ka, Kb;
                    if (prereq_device(&a->g, P, Q, EJ, EK, J, K, &Ja, &Jb, &Ka, &Kb))
                        for (int Kq = Ka; Kq <= Kb && ok; ++Kq) for (int Jq = Ja; Jq <= Jb; ++Jq) if (pf[Kq * Q->NJ + Jq] < need) { ok = 0; break; }
                }
                const int up = P->dk > 0 ? r - 1 : r + 1;                               /* the slab below in sweep direction */
                if (ok && K == 0 && up >= 0 && up < nslabs && !a->in_flag[q * (P->NJ + 1) + J]) ok = 0;   /* link_down */
                if (ok) { ++nready; rng = rng * 1664525u + 1013904223u; if ((rng >> 8) % (uint32_t)nready == 0) { pick_r = r; pick_t = t; } }
            }
        }
        if (pick_r < 0) { rc = -1; break; }
        Slab *a = &S[pick_r];
        int tk = a->taken[pick_t], q = 0; while (tk >= a->begin[q + 1]) ++q;
        const SP *P = &a->sp[q]; int J, K; ticket_to_JK(P, tk - a->begin[q], &J, &K);
        if (q >= 2 && a->completed[q - 2] != a->sp[q - 2].NJ * a->sp[q - 2].NK) { rc = -2; break; }
        for (int r2 = 0; r2 < nslabs; ++r2) if (r2 != pick_r && S[r2].ndone < S[r2].total) { int q2 = 0, t2 = S[r2].ndone; while (t2 >= S[r2].begin[q2 + 1]) ++q2; if (q2 != q) { ++overlap; break; } }
        for (int rk = P->rk_first + K * EK; rk < P->rk_first + (K + 1) * EK && rk <= P->rk_last; ++rk)
            for (int rj = 1 + J * EJ; rj < 1 + (J + 1) * EJ && rj < nj; ++rj)
                for (int ri = 1; ri < ni; ++ri)
                    relax_linked(a, q, P->di > 0 ? ri : ni - 1 - ri, P->dj > 0 ? rj : nj - 1 - rj, P->dk > 0 ? rk : nk - 1 - rk, P->di, P->dj, P->dk);
        /* hand-over: the last K block pushes its rows of the boundary plane (rk_last), row rj = 0 rides with J = 0 */
        const int down = P->dk > 0 ? pick_r + 1 : pick_r - 1;
        if (K == P->NK - 1 && down >= 0 && down < nslabs) {
            Slab *d = &S[down];
            const int kbnd = P->dk > 0 ? P->rk_last : nk - 1 - P->rk_last;
            for (int rj = (J == 0 ? 0 : 1 + J * EJ); rj < 1 + (J + 1) * EJ && rj < nj; ++rj) {
                const int j = P->dj > 0 ? rj : nj - 1 - rj;
                for (int i = 0; i < ni; ++i) {
                    const int64_t c = (int64_t)i + (int64_t)ni * (j + (int64_t)nj * kbnd);
                    d->in_phi[(int64_t)q * plane + i + (int64_t)ni * j] = phi[c];
                    d->in_tri[(int64_t)q * plane + i + (int64_t)ni * j] = ctri[c];
                }
            }
            d->in_flag[q * (P->NJ + 1) + J] = 1;
        }
        a->flags[(q & 1) * a->stride + K * P->NJ + J] = (uint32_t)(q + 1) * 2 + 1;
        ++a->completed[q]; ++a->ndone; ++all_done;
        a->taken[pick_t] = a->taken[--a->ntaken];
    }
    if (overlap_out) *overlap_out = overlap;
    for (int r = 0; r < nslabs; ++r) { free(S[r].flags); free(S[r].completed); free(S[r].taken); free(S[r].in_phi); free(S[r].in_tri); free(S[r].in_flag); }
    free(S);
    if (rc) return rc;
    if (g_poison_reads) return -4;
    return (memcmp(phi, ref_phi, sizeof(float) * V) || memcmp(ctri, ref_tri, 4 * V)) ? 1 : 0;
}
