/*
 * relax_emu.c -- CPU emulation of the relaxation sweep schedule (TEST INFRASTRUCTURE ONLY).
 *
 * sdfgen_b200/csrc/sdfb_sweep_relax.cu treats a Gauss-Seidel sweep (cpu_lib/makelevelset3.cpp:104-151) as the
 * fixed point of   new[v] = G(old[v], new[n_0(v)] .. new[n_6(v)])   and reaches it by chaotic iteration: round 0
 * evaluates every voxel against whatever its neighbours currently hold, later rounds re-evaluate the downstream
 * neighbours of the voxels that changed, always starting from the voxel's value at the START of the sweep.  This
 * file runs the same rules (candidate filter with the stamp memo, old-value log, revert, de-duplicated work
 * lists) on the CPU in a seeded RANDOM order inside every round -- the GPU's order is arbitrary too -- so that
 * tests/test_relax_emu.py can check the claim "any order ends in the serial result" against the serial oracle
 * without a GPU.  It cannot check memory-model behaviour; the GPU parity tests do that.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

float sdfo_point_triangle_distance(const float *x0, const float *x1, const float *x2, const float *x3);

#define TRI_MASK 0x07ffffffu
#define TRI_NONE 0x07ffffffu

static const int RDIRS[8][3] = { {+1,+1,+1}, {-1,-1,-1}, {+1,+1,-1}, {-1,-1,+1}, {+1,-1,+1}, {-1,+1,-1}, {+1,-1,-1}, {-1,+1,+1} };

static uint64_t rng_next(uint64_t *s) { *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17; return *s; }
static void shuffle(int64_t *a, int64_t n, uint64_t *s)
{
    for (int64_t i = n - 1; i > 0; --i) { int64_t j = (int64_t)(rng_next(s) % (uint64_t)(i + 1)); int64_t t = a[i]; a[i] = a[j]; a[j] = t; }
}

/*
 * cells_phi / cells_lo: (nkl+2) planes (halo plane first and last), i fastest, updated in place.
 * Returns the number of distance evaluations, or -1 on allocation failure.  *changed_out = net number of cells
 * whose triangle differs from the start of the sweep, *rounds_out = rounds until the work list ran empty.
 */
long sdfo_emu_sweep_relax(const uint32_t *tri, const float *x, float *cells_phi, uint32_t *cells_lo,
                          const float origin[3], float dx, int ni, int nj, int nk, int k_lo, int k_hi,
                          int sweep_index, uint64_t seed, long *changed_out, long *rounds_out)
{
    const int di = RDIRS[sweep_index % 8][0], dj = RDIRS[sweep_index % 8][1], dk = RDIRS[sweep_index % 8][2];
    const int64_t plane = (int64_t)ni * nj, ncell = plane * (k_hi - k_lo + 2);
    int ra = dk > 0 ? k_lo : nk - 1 - k_lo, rb = dk > 0 ? k_hi - 1 : nk - 1 - (k_hi - 1);
    int rk_first = ra < rb ? ra : rb, rk_last = ra < rb ? rb : ra;
    if (rk_first < 1) rk_first = 1;
    if (changed_out) *changed_out = 0;
    if (rounds_out) *rounds_out = 0;
    if (rk_first > rk_last || ni < 2 || nj < 2 || sweep_index + 1 >= 31) return 0;
    const uint32_t stamp = (uint32_t)(sweep_index + 1);
    uint8_t last[8][7];
    for (int c = 0; c < 8; ++c) for (int m = 0; m < 7; ++m) {
        last[c][m] = 0;
        int ci = (m == 0 || m == 2 || m == 4 || m == 6), cj = (m == 1 || m == 2 || m == 5 || m == 6), ck = (m >= 3);
        for (int e = sweep_index - 1; e >= 0; --e) {
            const int *d = RDIRS[e % 8];
            int same_i = d[0] == di, same_j = d[1] == dj, same_k = d[2] == dk;
            if ((!(ci || (c & 1)) || same_i) && (!(cj || (c & 2)) || same_j) && (!(ck || (c & 4)) || same_k)) { last[c][m] = (uint8_t)(e + 1); break; }
        }
    }
    float *old_phi = (float *)malloc(sizeof(float) * (size_t)ncell);
    uint32_t *old_lo = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)ncell);
    uint8_t *queued = (uint8_t *)calloc((size_t)ncell, 1);
    int64_t *list = (int64_t *)malloc(sizeof(int64_t) * (size_t)ncell), *next = (int64_t *)malloc(sizeof(int64_t) * (size_t)ncell);
    if (!old_phi || !old_lo || !queued || !list || !next) { free(old_phi); free(old_lo); free(queued); free(list); free(next); return -1; }
    const int64_t si = -(int64_t)di, sj = -(int64_t)dj * ni, sk = -(int64_t)dk * plane;
    const int64_t off[7] = { si, sj, si + sj, sk, si + sk, sj + sk, si + sj + sk };
    long evals = 0, net = 0, rounds = 0;
    uint64_t rs = seed ? seed : 0x9e3779b97f4a7c15ull;

    /* round 0: every voxel the sweep updates, in random order */
    int64_t n = 0;
    for (int rk = rk_first; rk <= rk_last; ++rk) for (int rj = 1; rj <= nj - 1; ++rj) for (int ri = 1; ri <= ni - 1; ++ri) {
        int i = di > 0 ? ri : ni - 1 - ri, j = dj > 0 ? rj : nj - 1 - rj, k = dk > 0 ? rk : nk - 1 - rk;
        list[n++] = (int64_t)i + (int64_t)ni * ((int64_t)j + (int64_t)nj * (int64_t)(k - k_lo + 1));
    }
    for (;;) {
        shuffle(list, n, &rs);
        int64_t nn = 0;
        for (int64_t q = 0; q < n; ++q) {
            const int64_t c = list[q];
            if (rounds > 0) queued[c] = 0;                     /* popped; may be scheduled again for the next round */
            const int64_t p = c / plane, rem = c - p * plane;
            const int j = (int)(rem / ni), i = (int)(rem - (int64_t)j * ni), k = (int)p - 1 + k_lo;
            const int ri = di > 0 ? i : ni - 1 - i, rj = dj > 0 ? j : nj - 1 - j, rk = dk > 0 ? k : nk - 1 - k;
            const int was_changed = (cells_lo[c] >> 27) == stamp;
            const float base_phi = was_changed ? old_phi[c] : cells_phi[c];
            const uint32_t base_lo = was_changed ? old_lo[c] : cells_lo[c];
            const int cls = (ri == ni - 1 ? 1 : 0) | (rj == nj - 1 ? 2 : 0) | (rk == nk - 1 ? 4 : 0);
            uint32_t nb[7], live = 0;
            for (int m = 0; m < 7; ++m) {
                nb[m] = cells_lo[c + off[m]];
                uint32_t thr = last[cls][m] ? ((uint32_t)last[cls][m] + 1u) << 27 : 0u;
                if (((nb[m] & TRI_MASK) != TRI_NONE) && (((nb[m] ^ base_lo) & TRI_MASK) != 0) && nb[m] >= thr) live |= 1u << m;
            }
            for (int m = 1; m < 7; ++m) for (int u = 0; u < m; ++u) if (((nb[u] ^ nb[m]) & TRI_MASK) == 0) live &= ~(1u << m);
            float phi = base_phi;
            uint32_t best = TRI_NONE;
            const float gx[3] = { (float)i * dx + origin[0], (float)j * dx + origin[1], (float)k * dx + origin[2] };
            for (int m = 0; m < 7; ++m) if ((live >> m) & 1u) {          /* the reference's order and strict "<" */
                uint32_t t = nb[m] & TRI_MASK;
                float d = sdfo_point_triangle_distance(gx, x + 3 * (size_t)tri[3 * (size_t)t], x + 3 * (size_t)tri[3 * (size_t)t + 1], x + 3 * (size_t)tri[3 * (size_t)t + 2]);
                ++evals;
                if (d < phi) { phi = d; best = t; }
            }
            const float new_phi = best != TRI_NONE ? phi : base_phi;
            const uint32_t new_lo = best != TRI_NONE ? ((stamp << 27) | best) : base_lo;
            if (new_lo != cells_lo[c] || memcmp(&new_phi, &cells_phi[c], 4) != 0) {
                if (!was_changed) { old_phi[c] = cells_phi[c]; old_lo[c] = cells_lo[c]; }
                cells_phi[c] = new_phi; cells_lo[c] = new_lo;
                net += (best != TRI_NONE ? 1 : 0) - (was_changed ? 1 : 0);
                for (int m = 1; m < 8; ++m) {                  /* the downstream neighbours this launch updates */
                    const int a = m & 1, b = (m >> 1) & 1, cc = (m >> 2) & 1;
                    if ((a && ri + 1 > ni - 1) || (b && rj + 1 > nj - 1) || (cc && rk + 1 > rk_last)) continue;
                    const int64_t d = c - (a ? si : 0) - (b ? sj : 0) - (cc ? sk : 0);
                    if (!queued[d]) { queued[d] = 1; next[nn++] = d; }
                }
            }
        }
        ++rounds;
        if (nn == 0) break;
        int64_t *t = list; list = next; next = t; n = nn;
    }
    free(old_phi); free(old_lo); free(queued); free(list); free(next);
    if (changed_out) *changed_out = net;
    if (rounds_out) *rounds_out = rounds;
    return evals;
}
