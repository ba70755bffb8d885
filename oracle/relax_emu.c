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

/* Design study ("what is next" for the rounds): with tile > 0 a push that stays inside the tile x tile x tile block
 * of the voxel that changed is processed in the SAME round (as a CTA that keeps the block's cells in shared memory could do);
 * only pushes that leave the block wait for the next round.  The result is the same -- any order is --; what changes is the
 * number of grid-wide rounds, which rounds_out then reports. */
static int emu_tile = 0;
void sdfo_emu_set_tile(int t) { emu_tile = t > 0 ? t : 0; }

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
/* round0: NULL = round 0 evaluates every voxel the sweep updates; otherwise one byte per cell, non-zero = evaluate in round 0
 * (the lookahead window of sdfb_sweep_relax.cu: sdfo_emu_look_scan / sdfo_emu_look_mark below).  change_log: NULL, or one
 * byte per cell that is set for every cell this sweep rewrites (the window's c_list). */
long sdfo_emu_sweep_relax_from(const uint32_t *tri, const float *x, float *cells_phi, uint32_t *cells_lo,
                               const float origin[3], float dx, int ni, int nj, int nk, int k_lo, int k_hi,
                               int sweep_index, uint64_t seed, const uint8_t *round0, uint8_t *change_log,
                               long *changed_out, long *rounds_out)
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
        const int64_t c0 = (int64_t)i + (int64_t)ni * ((int64_t)j + (int64_t)nj * (int64_t)(k - k_lo + 1));
        if (!round0 || round0[c0]) list[n++] = c0;
    }
    if (n == 0) { free(old_phi); free(old_lo); free(queued); free(list); free(next); return 0; }
    for (;;) {
        shuffle(list, n, &rs);
        int64_t nn = 0;
        for (int64_t q = 0; q < n; ++q) {
            const int64_t c = list[q];
            queued[c] = 0;                                     /* popped; may be scheduled again */
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
                if (!was_changed) { old_phi[c] = cells_phi[c]; old_lo[c] = cells_lo[c]; if (change_log) change_log[c] = 1; }
                cells_phi[c] = new_phi; cells_lo[c] = new_lo;
                net += (best != TRI_NONE ? 1 : 0) - (was_changed ? 1 : 0);
                for (int m = 1; m < 8; ++m) {                  /* the downstream neighbours this launch updates */
                    const int a = m & 1, b = (m >> 1) & 1, cc = (m >> 2) & 1;
                    if ((a && ri + 1 > ni - 1) || (b && rj + 1 > nj - 1) || (cc && rk + 1 > rk_last)) continue;
                    const int64_t d = c - (a ? si : 0) - (b ? sj : 0) - (cc ? sk : 0);
                    if (queued[d]) continue;
                    queued[d] = 1;
                    if (emu_tile > 0) {                        /* same block: this round (appended behind the entries still to come) */
                        const int i2 = i + (a ? di : 0), j2 = j + (b ? dj : 0), k2 = k + (cc ? dk : 0);
                        if (i2 / emu_tile == i / emu_tile && j2 / emu_tile == j / emu_tile && k2 / emu_tile == k / emu_tile && n < ncell) { list[n++] = d; continue; }
                    }
                    next[nn++] = d;
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

long sdfo_emu_sweep_relax(const uint32_t *tri, const float *x, float *cells_phi, uint32_t *cells_lo,
                          const float origin[3], float dx, int ni, int nj, int nk, int k_lo, int k_hi,
                          int sweep_index, uint64_t seed, long *changed_out, long *rounds_out)
{
    return sdfo_emu_sweep_relax_from(tri, x, cells_phi, cells_lo, origin, dx, ni, nj, nk, k_lo, k_hi, sweep_index, seed,
                                     NULL, NULL, changed_out, rounds_out);
}

/* ---- lookahead window (k_look_scan / k_look_mark of sdfb_sweep_relax.cu), whole grids only ------------------------------
 * sdfo_emu_look_scan: on the cells as they are BEFORE sweep s_lo, find for every sweep s of [s_lo, s_hi) (at most 8) the
 * voxels one of whose candidates -- the neighbours sweep s would evaluate if nothing around the voxel changed until then,
 * by the memo table of sweep s -- beats the voxel's distance: marks[(s - s_lo) * ncell + c] = 1.
 * Interior voxels of a window that starts at a multiple of 8 walk the 26 neighbour offsets once, in the order in which the
 * window's sweeps first examine them, and drop a candidate whose triangle an earlier offset within Manhattan distance
 * `dedupe` already names (dedupe = 0: the sweeps' own rule, the earlier neighbours of the same sweep); all other voxels
 * evaluate each sweep's seven neighbours with the sweep's rule.  Returns the number of distance evaluations. */
static void emu_last_table(int sweep_index, uint8_t last[8][7])
{
    const int di = RDIRS[sweep_index % 8][0], dj = RDIRS[sweep_index % 8][1], dk = RDIRS[sweep_index % 8][2];
    for (int c = 0; c < 8; ++c) for (int m = 0; m < 7; ++m) {
        last[c][m] = 0;
        int ci = (m == 0 || m == 2 || m == 4 || m == 6), cj = (m == 1 || m == 2 || m == 5 || m == 6), ck = (m >= 3);
        for (int e = sweep_index - 1; e >= 0; --e) {
            const int *d = RDIRS[e % 8];
            int same_i = d[0] == di, same_j = d[1] == dj, same_k = d[2] == dk;
            if ((!(ci || (c & 1)) || same_i) && (!(cj || (c & 2)) || same_j) && (!(ck || (c & 4)) || same_k)) { last[c][m] = (uint8_t)(e + 1); break; }
        }
    }
}

long sdfo_emu_look_scan(const uint32_t *tri, const float *x, const float *cells_phi, const uint32_t *cells_lo,
                        const float origin[3], float dx, int ni, int nj, int nk, int s_lo, int s_hi, int dedupe, uint8_t *marks)
{
    const int64_t plane = (int64_t)ni * nj, ncell = plane * (nk + 2);
    const int ns = s_hi - s_lo;
    if (ns < 1 || ns > 8 || ni < 2 || nj < 2 || nk < 2) return -1;
    uint8_t last[8][8][7];
    for (int w = 0; w < ns; ++w) emu_last_table(s_lo + w, last[w]);
    /* window order of the 26 offsets (first examination), as kLookOrder: {oi, oj, ok, w, m} */
    int order[26][5], nord = 0;
    const int standard = (s_lo % 8) == 0;
    if (standard) {
        uint8_t seen[27] = {0};
        for (int q = 0; q < 8; ++q) for (int m = 0; m < 7; ++m) {
            const int *d = RDIRS[q];
            int oi = -d[0] * (m == 0 || m == 2 || m == 4 || m == 6), oj = -d[1] * (m == 1 || m == 2 || m == 5 || m == 6), ok = -d[2] * (m >= 3);
            int id = (ok + 1) * 9 + (oj + 1) * 3 + (oi + 1);
            if (seen[id]) continue;
            seen[id] = 1;
            order[nord][0] = oi; order[nord][1] = oj; order[nord][2] = ok; order[nord][3] = q; order[nord][4] = m; ++nord;
        }
    }
    long evals = 0;
    for (int k = 0; k < nk; ++k) for (int j = 0; j < nj; ++j) for (int i = 0; i < ni; ++i) {
        const int64_t c = (int64_t)i + (int64_t)ni * ((int64_t)j + (int64_t)nj * (int64_t)(k + 1));
        const uint32_t own = cells_lo[c];
        const float phi = cells_phi[c];
        const float gx[3] = { (float)i * dx + origin[0], (float)j * dx + origin[1], (float)k * dx + origin[2] };
        const int interior = i >= 1 && i <= ni - 2 && j >= 1 && j <= nj - 2 && k >= 1 && k <= nk - 2;
        if (standard && interior) {
            uint32_t word[26];
            for (int n = 0; n < 26; ++n) {
                const int *o = order[n];
                word[n] = cells_lo[c + o[0] + (int64_t)o[1] * ni + (int64_t)o[2] * plane];
                const int w = o[3];                               /* standard window: position in the window = direction */
                if (w >= ns) continue;
                const uint32_t thr = last[w][0][o[4]] ? ((uint32_t)last[w][0][o[4]] + 1u) << 27 : 0u;
                if (((word[n] & TRI_MASK) == TRI_NONE) || (((word[n] ^ own) & TRI_MASK) == 0) || word[n] < thr) continue;
                int dup = 0;
                for (int u = 0; u < n && !dup; ++u) {
                    const int *b = order[u];
                    int partner;
                    if (dedupe > 0) partner = abs(o[0] - b[0]) + abs(o[1] - b[1]) + abs(o[2] - b[2]) <= dedupe;
                    else {
                        partner = 0;
                        for (int up = 0; up < o[4]; ++up) {
                            const int *d = RDIRS[w];
                            int pi = -d[0] * (up == 0 || up == 2 || up == 4 || up == 6), pj = -d[1] * (up == 1 || up == 2 || up == 5 || up == 6), pk = -d[2] * (up >= 3);
                            if (b[0] == pi && b[1] == pj && b[2] == pk) partner = 1;
                        }
                    }
                    if (partner && ((word[u] ^ word[n]) & TRI_MASK) == 0) dup = 1;
                }
                if (dup) continue;
                const uint32_t t = word[n] & TRI_MASK;
                const float d = sdfo_point_triangle_distance(gx, x + 3 * (size_t)tri[3 * (size_t)t], x + 3 * (size_t)tri[3 * (size_t)t + 1], x + 3 * (size_t)tri[3 * (size_t)t + 2]);
                ++evals;
                if (d < phi) marks[(int64_t)w * ncell + c] = 1;
            }
            continue;
        }
        for (int w = 0; w < ns; ++w) {
            const int *d = RDIRS[(s_lo + w) % 8];
            const int ri = d[0] > 0 ? i : ni - 1 - i, rj = d[1] > 0 ? j : nj - 1 - j, rk = d[2] > 0 ? k : nk - 1 - k;
            if (ri < 1 || rj < 1 || rk < 1) continue;             /* on the face the sweep starts from: never updated by it */
            const int cls = (ri == ni - 1 ? 1 : 0) | (rj == nj - 1 ? 2 : 0) | (rk == nk - 1 ? 4 : 0);
            const int64_t si = -(int64_t)d[0], sj = -(int64_t)d[1] * ni, sk = -(int64_t)d[2] * plane;
            const int64_t off[7] = { si, sj, si + sj, sk, si + sk, sj + sk, si + sj + sk };
            uint32_t nb[7], live = 0;
            for (int m = 0; m < 7; ++m) {
                nb[m] = cells_lo[c + off[m]];
                const uint32_t thr = last[w][cls][m] ? ((uint32_t)last[w][cls][m] + 1u) << 27 : 0u;
                if (((nb[m] & TRI_MASK) != TRI_NONE) && (((nb[m] ^ own) & TRI_MASK) != 0) && nb[m] >= thr) live |= 1u << m;
            }
            for (int m = 1; m < 7; ++m) for (int u = 0; u < m; ++u) if (((nb[u] ^ nb[m]) & TRI_MASK) == 0) live &= ~(1u << m);
            for (int m = 0; m < 7; ++m) if ((live >> m) & 1u) {
                const uint32_t t = nb[m] & TRI_MASK;
                const float dd = sdfo_point_triangle_distance(gx, x + 3 * (size_t)tri[3 * (size_t)t], x + 3 * (size_t)tri[3 * (size_t)t + 1], x + 3 * (size_t)tri[3 * (size_t)t + 2]);
                ++evals;
                if (dd < phi) marks[(int64_t)w * ncell + c] = 1;
            }
        }
    }
    return evals;
}

/* sdfo_emu_look_mark: round-0 set of sweep `sweep_index` = its marks  u  every logged cell and its seven downstream
 * neighbours, as far as the sweep updates them.  round0 must be zeroed by the caller.  Returns the number of cells set. */
long sdfo_emu_look_mark(const uint8_t *marks_of_sweep, const uint8_t *change_log, int ni, int nj, int nk, int sweep_index, uint8_t *round0)
{
    const int *d = RDIRS[sweep_index % 8];
    const int64_t plane = (int64_t)ni * nj, ncell = plane * (nk + 2);
    long n = 0;
    for (int64_t c = 0; c < ncell; ++c) if (marks_of_sweep[c] && !round0[c]) { round0[c] = 1; ++n; }
    for (int k = 0; k < nk; ++k) for (int j = 0; j < nj; ++j) for (int i = 0; i < ni; ++i) {
        const int64_t c = (int64_t)i + (int64_t)ni * ((int64_t)j + (int64_t)nj * (int64_t)(k + 1));
        if (!change_log[c]) continue;
        const int ri = d[0] > 0 ? i : ni - 1 - i, rj = d[1] > 0 ? j : nj - 1 - j, rk = d[2] > 0 ? k : nk - 1 - k;
        for (int m = 0; m < 8; ++m) {
            const int a = ri + (m & 1), b = rj + ((m >> 1) & 1), cc = rk + ((m >> 2) & 1);
            if (a < 1 || a > ni - 1 || b < 1 || b > nj - 1 || cc < 1 || cc > nk - 1) continue;
            const int ii = d[0] > 0 ? a : ni - 1 - a, jj = d[1] > 0 ? b : nj - 1 - b, kk = d[2] > 0 ? cc : nk - 1 - cc;
            const int64_t t = (int64_t)ii + (int64_t)ni * ((int64_t)jj + (int64_t)nj * (int64_t)(kk + 1));
            if (!round0[t]) { round0[t] = 1; ++n; }
        }
    }
    return n;
}
