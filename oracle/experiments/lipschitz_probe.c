/*
 * lipschitz_probe.c -- CPU experiment (TEST INFRASTRUCTURE / design study, never linked into the product).
 *
 * SURVEY.md section 7 proposes a conservative 1-Lipschitz prune for the sweeps: the candidate triangle t of neighbour n
 * satisfies d(v, t) >= d(n, t) - |v - n| = phi(n) - |delta| dx, so it cannot win at voxel v when
 * phi(n) - |delta| dx >= phi(v) (plus a rounding margin).  This program runs the reference's 16 serial sweeps
 * (cpu_lib/makelevelset3.cpp:104-151, order of :245-248) with the pruning the CUDA schedules already do -- own triangle,
 * duplicates among the 7 neighbours, stamp memo -- and counts, per sweep, how many of the REMAINING evaluations the bound
 * would remove, and whether any removed candidate would in fact have won (exactness violations; must be 0).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

float sdfo_point_triangle_distance(const float *x0, const float *x1, const float *x2, const float *x3);
static const int DIRS[8][3] = { {+1,+1,+1}, {-1,-1,-1}, {+1,+1,-1}, {-1,-1,+1}, {+1,-1,+1}, {-1,+1,-1}, {+1,-1,-1}, {-1,+1,+1} };
static const int OFF[7][3] = { {1,0,0}, {0,1,0}, {1,1,0}, {0,0,1}, {1,0,1}, {0,1,1}, {1,1,1} };

/* out[s*4 + 0] evaluations after the existing pruning, +1 of those removed by the bound, +2 violations, +3 changed cells */
int lipschitz_probe(const uint32_t *tri, const float *x, float *phi, int32_t *ctri, const float origin[3], float dx,
                    int ni, int nj, int nk, int nsweeps, double margin_rel, int64_t *out)
{
    const int64_t V = (int64_t)ni * nj * nk;
    uint8_t *stamp = calloc(V, 1);
    if (!stamp) return -1;
    const float dl[7] = { 1.f, 1.f, sqrtf(2.f), 1.f, sqrtf(2.f), sqrtf(2.f), sqrtf(3.f) };
    for (int s = 0; s < nsweeps; ++s) {
        const int di = DIRS[s % 8][0], dj = DIRS[s % 8][1], dk = DIRS[s % 8][2];
        uint8_t last[8][7];
        for (int c = 0; c < 8; ++c) for (int m = 0; m < 7; ++m) {
            last[c][m] = 0;
            int ci = (m == 0 || m == 2 || m == 4 || m == 6), cj = (m == 1 || m == 2 || m == 5 || m == 6), ck = (m >= 3);
            for (int e = s - 1; e >= 0; --e) {
                const int *d = DIRS[e % 8];
                if ((!(ci || (c & 1)) || d[0] == di) && (!(cj || (c & 2)) || d[1] == dj) && (!(ck || (c & 4)) || d[2] == dk)) { last[c][m] = (uint8_t)(e + 1); break; }
            }
        }
        int64_t evals = 0, pruned = 0, viol = 0, changed = 0;
        for (int rk = 1; rk < nk; ++rk) for (int rj = 1; rj < nj; ++rj) for (int ri = 1; ri < ni; ++ri) {
            const int i = di > 0 ? ri : ni - 1 - ri, j = dj > 0 ? rj : nj - 1 - rj, k = dk > 0 ? rk : nk - 1 - rk;
            const int cls = (ri == ni - 1 ? 1 : 0) | (rj == nj - 1 ? 2 : 0) | (rk == nk - 1 ? 4 : 0);
            const int64_t c0 = (int64_t)i + (int64_t)ni * (j + (int64_t)nj * k);
            const float gx[3] = { i * dx + origin[0], j * dx + origin[1], k * dx + origin[2] };
            int32_t seen[7]; int ns = 0, did = 0;
            const int32_t own0 = ctri[c0];
            for (int m = 0; m < 7; ++m) {
                const int64_t c1 = (int64_t)(i - di * OFF[m][0]) + (int64_t)ni * ((j - dj * OFF[m][1]) + (int64_t)nj * (k - dk * OFF[m][2]));
                const int32_t t = ctri[c1];
                if (t < 0) continue;
                int dup = 0;
                for (int u = 0; u < ns; ++u) dup |= seen[u] == t;
                seen[ns++] = t;
                if (dup || t == own0) continue;
                if (stamp[c1] <= last[cls][m] && last[cls][m] != 0) continue;        /* memo: unchanged since this voxel last looked */
                ++evals;
                const float bound = phi[c1] - dl[m] * dx;
                const int prune = (double)bound - margin_rel * ((double)phi[c1] + (double)dx) >= (double)phi[c0];
                const uint32_t *tv = tri + 3 * (size_t)t;
                const float d = sdfo_point_triangle_distance(gx, x + 3 * (size_t)tv[0], x + 3 * (size_t)tv[1], x + 3 * (size_t)tv[2]);
                if (prune) { ++pruned; if (d < phi[c0]) ++viol; }
                if (d < phi[c0]) { phi[c0] = d; ctri[c0] = t; did = 1; }
            }
            if (did) { stamp[c0] = (uint8_t)(s + 1); ++changed; }
        }
        out[s * 4 + 0] = evals; out[s * 4 + 1] = pruned; out[s * 4 + 2] = viol; out[s * 4 + 3] = changed;
    }
    free(stamp);
    return 0;
}
