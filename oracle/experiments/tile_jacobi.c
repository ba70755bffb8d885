/*
 * tile_jacobi.c -- CPU experiment (TEST INFRASTRUCTURE / design study, never linked into the product).
 *
 * Question for the next round: can a first-pass sweep be run tile-parallel instead of as one 1534-level wavefront?
 * A Gauss-Seidel sweep (cpu_lib/makelevelset3.cpp:104-151) is the unique solution of
 *     new[v] = G(old[v], new[n_0(v)] .. new[n_6(v)])          (upstream neighbours, fixed order, strict <)
 * so it can be reached by iterating over TILES: in iteration r every "dirty" tile is swept serially from its
 * start-of-sweep values against the halo cells its upstream neighbour tiles held after iteration r-1 (Jacobi between
 * tiles, Gauss-Seidel inside); a tile is dirty in iteration r+1 if one of the halo cells it reads changed in
 * iteration r.  All tiles of an iteration are independent.  This program runs that scheme for the 16 sweeps, checks
 * after every sweep that the result equals the serial sweep bit for bit, and reports per sweep: iterations, tile sweeps
 * (work amplification = tile sweeps / tiles) and distance evaluations.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

float sdfo_point_triangle_distance(const float *x0, const float *x1, const float *x2, const float *x3);

static const int DIRS[8][3] = { {+1,+1,+1}, {-1,-1,-1}, {+1,+1,-1}, {-1,-1,+1}, {+1,-1,+1}, {-1,+1,-1}, {+1,-1,-1}, {-1,+1,+1} };

typedef struct { int ni, nj, nk, di, dj, dk; float dx, o[3]; const uint32_t *tri; const float *x; } Ctx;

static inline int64_t idx(const Ctx *c, int ri, int rj, int rk)      /* sweep-relative -> linear (i fastest) */
{
    int i = c->di > 0 ? ri : c->ni - 1 - ri, j = c->dj > 0 ? rj : c->nj - 1 - rj, k = c->dk > 0 ? rk : c->nk - 1 - rk;
    return (int64_t)i + (int64_t)c->ni * ((int64_t)j + (int64_t)c->nj * k);
}

/* one voxel: G(start-of-sweep value, current upstream neighbours); reads rd_*, writes wr_* */
/* incremental-work counter: voxels of re-sweeps (iteration >= 1) with at least one input that differs from the
 * previous iteration -- what an incremental (relaxation-style) re-sweep would have to recompute */
static long g_reeval;
long tile_jacobi_last_reeval(void) { return g_reeval; }

static inline long relax_voxel(const Ctx *c, int ri, int rj, int rk, const float *old_phi, const int32_t *old_tri,
                               const float *rd_phi, const int32_t *rd_tri, float *wr_phi, int32_t *wr_tri,
                               int bi, int bj, int bk)   /* tile origin: neighbours inside the tile come from wr_* */
{
    const int64_t c0 = idx(c, ri, rj, rk);
    int i = c->di > 0 ? ri : c->ni - 1 - ri, j = c->dj > 0 ? rj : c->nj - 1 - rj, k = c->dk > 0 ? rk : c->nk - 1 - rk;
    float gx[3] = { i * c->dx + c->o[0], j * c->dx + c->o[1], k * c->dx + c->o[2] };
    float phi = old_phi[c0]; int32_t best = old_tri[c0];
    static const int OFF[7][3] = { {1,0,0}, {0,1,0}, {1,1,0}, {0,0,1}, {1,0,1}, {0,1,1}, {1,1,1} };   /* :143-149 */
    long ev = 0;
    for (int m = 0; m < 7; ++m) {
        int ui = ri - OFF[m][0], uj = rj - OFF[m][1], uk = rk - OFF[m][2];
        int64_t c1 = idx(c, ui, uj, uk);
        int inside = ui >= bi && uj >= bj && uk >= bk;                 /* same tile -> this iteration's value */
        int32_t t = inside ? wr_tri[c1] : rd_tri[c1];
        if (t >= 0) {
            const uint32_t *tv = c->tri + 3 * (size_t)t;
            float d = sdfo_point_triangle_distance(gx, c->x + 3 * (size_t)tv[0], c->x + 3 * (size_t)tv[1], c->x + 3 * (size_t)tv[2]);
            ++ev;
            if (d < phi) { phi = d; best = t; }
        }
    }
    wr_phi[c0] = phi; wr_tri[c0] = best;
    return ev;
}

/*
 * phi/tri: state, updated in place through the 16 sweeps.  B: tile edge.  report[s*4+0..3] = iterations, tile sweeps,
 * tiles, evaluations of sweep s; dirty_hist[s*64 + r] = dirty tiles in iteration r (r < 64).
 * Returns 0 if every sweep matched the serial sweep bit for bit, else 1 + index of the first sweep that did not.
 */
int tile_jacobi_run(const uint32_t *tri, const float *x, float *phi, int32_t *ctri, const float origin[3], float dx,
                    int ni, int nj, int nk, int B, int nsweeps, long *report, long *dirty_hist, long *reeval_out)
{
    const int64_t V = (int64_t)ni * nj * nk;
    float *old_phi = malloc(sizeof(float) * V), *a_phi = malloc(sizeof(float) * V), *b_phi = malloc(sizeof(float) * V), *s_phi = malloc(sizeof(float) * V);
    int32_t *old_tri = malloc(4 * V), *a_tri = malloc(4 * V), *b_tri = malloc(4 * V), *s_tri = malloc(4 * V);
    const int TI = (ni - 1 + B - 1) / B, TJ = (nj - 1 + B - 1) / B, TK = (nk - 1 + B - 1) / B, NT = TI * TJ * TK;
    uint8_t *dirty = malloc(NT), *next_dirty = malloc(NT);
    float *p_phi = malloc(sizeof(float) * V); int32_t *p_tri = malloc(4 * V);      /* state before iteration r-1 */
    int bad = 0;
    for (int s = 0; s < nsweeps && !bad; ++s) {
        Ctx c = { ni, nj, nk, DIRS[s % 8][0], DIRS[s % 8][1], DIRS[s % 8][2], dx, { origin[0], origin[1], origin[2] }, tri, x };
        memcpy(old_phi, phi, sizeof(float) * V); memcpy(old_tri, ctri, 4 * V);
        /* serial sweep for the check */
        memcpy(s_phi, phi, sizeof(float) * V); memcpy(s_tri, ctri, 4 * V);
        for (int rk = 1; rk < nk; ++rk) for (int rj = 1; rj < nj; ++rj) for (int ri = 1; ri < ni; ++ri)
            relax_voxel(&c, ri, rj, rk, s_phi, s_tri, s_phi, s_tri, s_phi, s_tri, 0, 0, 0);
        /* tile iteration: a = state after iteration r-1, b = after iteration r */
        memcpy(a_phi, phi, sizeof(float) * V); memcpy(a_tri, ctri, 4 * V);
        memcpy(b_phi, phi, sizeof(float) * V); memcpy(b_tri, ctri, 4 * V);
        memcpy(p_phi, phi, sizeof(float) * V); memcpy(p_tri, ctri, 4 * V);
        memset(dirty, 1, NT);
        g_reeval = 0;
        long iters = 0, tsweeps = 0, evals = 0;
        for (;;) {
            long nd = 0;
            for (int t = 0; t < NT; ++t) nd += dirty[t];
            if (nd == 0) break;
            if (iters < 64) dirty_hist[s * 64 + iters] = nd;
            ++iters; tsweeps += nd;
            #pragma omp parallel for schedule(dynamic, 1) reduction(+:evals)
            for (int t = 0; t < NT; ++t) {
                if (!dirty[t]) continue;
                int I = t % TI, J = (t / TI) % TJ, K = t / (TI * TJ);
                int bi = 1 + I * B, bj = 1 + J * B, bk = 1 + K * B;
                int ei = bi + B < ni ? bi + B : ni, ej = bj + B < nj ? bj + B : nj, ek = bk + B < nk ? bk + B : nk;
                long re = 0;
                for (int rk = bk; rk < ek; ++rk) for (int rj = bj; rj < ej; ++rj) for (int ri = bi; ri < ei; ++ri) {
                    evals += relax_voxel(&c, ri, rj, rk, old_phi, old_tri, a_phi, a_tri, b_phi, b_tri, bi, bj, bk);
                    if (iters > 1) {                       /* iters was already incremented: this is iteration iters-1 >= 1 */
                        static const int OFF[7][3] = { {1,0,0}, {0,1,0}, {1,1,0}, {0,0,1}, {1,0,1}, {0,1,1}, {1,1,1} };
                        int ch = 0;
                        for (int m = 0; m < 7 && !ch; ++m) {
                            int ui = ri - OFF[m][0], uj = rj - OFF[m][1], uk = rk - OFF[m][2];
                            int64_t q = idx(&c, ui, uj, uk);
                            int inside = ui >= bi && uj >= bj && uk >= bk;
                            /* input now: b (inside) or a (halo, = previous iteration's output); input then: a (inside) or p (halo) */
                            if (inside) ch = (a_tri[q] != b_tri[q]) || memcmp(&a_phi[q], &b_phi[q], 4);
                            else ch = (p_tri[q] != a_tri[q]) || memcmp(&p_phi[q], &a_phi[q], 4);
                        }
                        re += ch;
                    }
                }
                #pragma omp atomic
                g_reeval += re;
            }
            /* which tiles read a halo cell that changed in this iteration? */
            #pragma omp parallel for schedule(dynamic, 4)
            for (int t = 0; t < NT; ++t) {
                int I = t % TI, J = (t / TI) % TJ, K = t / (TI * TJ);
                int bi = 1 + I * B, bj = 1 + J * B, bk = 1 + K * B;
                int ei = bi + B < ni ? bi + B : ni, ej = bj + B < nj ? bj + B : nj, ek = bk + B < nk ? bk + B : nk;
                int d = 0;
                for (int rk = bk - 1; rk < ek && !d; ++rk) for (int rj = bj - 1; rj < ej && !d; ++rj) for (int ri = bi - 1; ri < ei; ++ri) {
                    if (ri >= bi && rj >= bj && rk >= bk) { ri = ei; continue; }            /* interior: skip the row's rest */
                    int64_t q = idx(&c, ri, rj, rk);
                    if (a_tri[q] != b_tri[q] || memcmp(&a_phi[q], &b_phi[q], 4)) { d = 1; break; }
                }
                next_dirty[t] = (uint8_t)d;
            }
            /* p <- a, a <- b */
            memcpy(p_phi, a_phi, sizeof(float) * V); memcpy(p_tri, a_tri, 4 * V);
            memcpy(a_phi, b_phi, sizeof(float) * V); memcpy(a_tri, b_tri, 4 * V);
            memcpy(dirty, next_dirty, NT);
        }
        report[s * 4 + 0] = iters; report[s * 4 + 1] = tsweeps; report[s * 4 + 2] = NT; report[s * 4 + 3] = evals;
        reeval_out[s] = g_reeval;
        if (memcmp(b_phi, s_phi, sizeof(float) * V) || memcmp(b_tri, s_tri, 4 * V)) bad = 1 + s;
        memcpy(phi, b_phi, sizeof(float) * V); memcpy(ctri, b_tri, 4 * V);
    }
    free(old_phi); free(a_phi); free(b_phi); free(s_phi); free(old_tri); free(a_tri); free(b_tri); free(s_tri); free(dirty); free(next_dirty); free(p_phi); free(p_tri);
    return bad;
}
