/*
 * sdf_oracle.c -- CPU restatement of SDFGenFast's make_level_set3 (TEST INFRASTRUCTURE ONLY).
 *
 * This file is the parity oracle for the CUDA path in sdfgen_b200/csrc.  It is test
 * infrastructure: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it.  The product never routes through it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement bit-for-bit against
 *   (1) the reference's own compiled CPU code (oracle/_ref, built by oracle/Makefile from
 *       /root/reference/cpu_lib/makelevelset3.cpp, single-threaded), when it is present, and
 *   (2) golden fixtures under tests/golden/ that were produced by that compiled reference
 *       (tests/golden/make_golden.py), including the survey's known-answer sha256 of the
 *       64x85x105 test-mesh .sdf.
 *
 * What it restates (all citations are /root/reference/ file:line):
 *   point_segment_distance   cpu_lib/makelevelset3.cpp:21-34
 *   point_triangle_distance  cpu_lib/makelevelset3.cpp:49-70
 *   check_neighbour          cpu_lib/makelevelset3.cpp:90-102
 *   sweep (serial)           cpu_lib/makelevelset3.cpp:104-127
 *   orientation              cpu_lib/makelevelset3.cpp:155-165
 *   point_in_triangle_2d     cpu_lib/makelevelset3.cpp:169-187
 *   make_level_set3 driver   cpu_lib/makelevelset3.cpp:192-304 (num_threads=1 semantics: the
 *                            multi-threaded reference is racy)
 *   vector helpers           common/vec.h:216-223 (mag2), :240-255 (dist2/dist), :377-383 (dot)
 *   min/max/clamp            common/util.h:22-23 (std::min/max), :59-61, :113-115, :341-347
 *   grid layout              common/array3.h:111-115  (i fastest: a[i + ni*(j + nj*k)])
 *
 * Differences from the reference on purpose: plain C, 64-bit linear indices (the reference's
 * int index overflows at 2^31 voxels, common/array3.h:59-61), staged outputs (band-only phi /
 * closest_tri, intersection counts) because parity is graded on them and the reference keeps
 * them as locals (cpu_lib/makelevelset3.cpp:198-199), optional k-slab window for phases A/C,
 * and evaluation counters used for DESIGN.md.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (no -march=native: FMA contraction would
 * change results).  See oracle/Makefile.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__FAST_MATH__)
#error "the oracle must not be built with -ffast-math"
#endif

typedef struct { float v[3]; } v3;

/* ---- small float helpers, same operation order as common/vec.h ---- */
static inline v3 v3sub(v3 a, v3 b) { v3 r = {{a.v[0]-b.v[0], a.v[1]-b.v[1], a.v[2]-b.v[2]}}; return r; }
/* common/vec.h:377-383  d=a0*b0; d+=a1*b1; d+=a2*b2 */
static inline float v3dot(v3 a, v3 b) { float d = a.v[0]*b.v[0]; d += a.v[1]*b.v[1]; d += a.v[2]*b.v[2]; return d; }
/* common/vec.h:216-223 */
static inline float v3mag2(v3 a) { float l = a.v[0]*a.v[0]; l += a.v[1]*a.v[1]; l += a.v[2]*a.v[2]; return l; }
/* common/vec.h:240-255 */
static inline float v3dist(v3 a, v3 b)
{
    float t0 = a.v[0]-b.v[0], t1 = a.v[1]-b.v[1], t2 = a.v[2]-b.v[2];
    float d = t0*t0; d += t1*t1; d += t2*t2;
    return sqrtf(d);
}
/* std::min / std::max semantics (NaN-order-sensitive): min(a,b) = (b<a)?b:a ; max(a,b) = (a<b)?b:a */
static inline float  fmin_std(float a, float b)   { return (b < a) ? b : a; }
static inline float  fmax_std(float a, float b)   { return (a < b) ? b : a; }
static inline double dmin_std(double a, double b) { return (b < a) ? b : a; }
static inline double dmax_std(double a, double b) { return (a < b) ? b : a; }
/* common/util.h:59-61 and :113-115: min(a1, min(a2,a3)) */
static inline double dmin3(double a, double b, double c) { return dmin_std(a, dmin_std(b, c)); }
static inline double dmax3(double a, double b, double c) { return dmax_std(a, dmax_std(b, c)); }
/* common/util.h:341-347 */
static inline int iclamp(int a, int lo, int hi) { if (a < lo) return lo; else if (a > hi) return hi; else return a; }

/* cpu_lib/makelevelset3.cpp:21-34 */
static float seg_distance(v3 x0, v3 x1, v3 x2)
{
    v3 e = v3sub(x2, x1);
    double m2 = v3mag2(e);                       /* float result widened to double (:24) */
    float s12 = (float)(v3dot(v3sub(x2, x0), e) / m2);  /* double divide, narrowed (:26) */
    if (s12 < 0) s12 = 0; else if (s12 > 1) s12 = 1;
    float om = 1 - s12;
    /* s12*x1 + (1-s12)*x2, component-wise (:33, common/vec.h:125-137,331-337) */
    v3 p = {{ s12*x1.v[0] + om*x2.v[0], s12*x1.v[1] + om*x2.v[1], s12*x1.v[2] + om*x2.v[2] }};
    return v3dist(x0, p);
}

/* cpu_lib/makelevelset3.cpp:49-70 */
float sdfo_point_triangle_distance(const float *px0, const float *px1, const float *px2, const float *px3)
{
    v3 x0, x1, x2, x3;
    memcpy(&x0, px0, 12); memcpy(&x1, px1, 12); memcpy(&x2, px2, 12); memcpy(&x3, px3, 12);
    v3 x13 = v3sub(x1, x3), x23 = v3sub(x2, x3), x03 = v3sub(x0, x3);
    float m13 = v3mag2(x13), m23 = v3mag2(x23), d = v3dot(x13, x23);
    float invdet = 1.f / fmax_std(m13*m23 - d*d, 1e-30f);
    float a = v3dot(x13, x03), b = v3dot(x23, x03);
    float w23 = invdet*(m23*a - d*b);
    float w31 = invdet*(m13*b - d*a);
    float w12 = 1 - w23 - w31;
    if (w23 >= 0 && w31 >= 0 && w12 >= 0) {
        /* w23*x1 + w31*x2 + w12*x3 evaluated left to right (:61) */
        v3 p = {{ w23*x1.v[0] + w31*x2.v[0] + w12*x3.v[0],
                  w23*x1.v[1] + w31*x2.v[1] + w12*x3.v[1],
                  w23*x1.v[2] + w31*x2.v[2] + w12*x3.v[2] }};
        return v3dist(x0, p);
    } else if (w23 > 0) {
        return fmin_std(seg_distance(x0, x1, x2), seg_distance(x0, x1, x3));
    } else if (w31 > 0) {
        return fmin_std(seg_distance(x0, x1, x2), seg_distance(x0, x2, x3));
    } else {
        return fmin_std(seg_distance(x0, x1, x3), seg_distance(x0, x2, x3));
    }
}

/* cpu_lib/makelevelset3.cpp:155-165 */
static int orient2d(double x1, double y1, double x2, double y2, double *twice_area)
{
    *twice_area = y1*x2 - x1*y2;
    if (*twice_area > 0) return 1;
    else if (*twice_area < 0) return -1;
    else if (y2 > y1) return 1;
    else if (y2 < y1) return -1;
    else if (x1 > x2) return 1;
    else if (x1 < x2) return -1;
    else return 0;
}

/* cpu_lib/makelevelset3.cpp:169-187 */
static int in_triangle_2d(double x0, double y0, double x1, double y1, double x2, double y2,
                          double x3, double y3, double *a, double *b, double *c)
{
    x1 -= x0; x2 -= x0; x3 -= x0;
    y1 -= y0; y2 -= y0; y3 -= y0;
    int sa = orient2d(x2, y2, x3, y3, a);
    if (sa == 0) return 0;
    int sb = orient2d(x3, y3, x1, y1, b);
    if (sb != sa) return 0;
    int sc = orient2d(x1, y1, x2, y2, c);
    if (sc != sa) return 0;
    double sum = *a + *b + *c;
    *a /= sum; *b /= sum; *c /= sum;
    return 1;
}

/* ------------------------------------------------------------------------------------------ */

typedef struct {
    /* nullable snapshot/debug outputs; each V = ni*nj*(k_hi-k_lo) elements, i fastest */
    float   *phi_band;     /* unsigned phi after phase A */
    int32_t *tri_band;     /* closest_tri after phase A */
    int32_t *counts;       /* intersection_count after phase A */
    float   *phi_swept;    /* unsigned phi after the sweeps */
    int32_t *tri_final;    /* closest_tri after the sweeps */
    /* counters (nullable): [0]=band evals, [1..16]=check_neighbour evals per sweep,
       [17..32]=updates per sweep, [33..48]=evals left per sweep if a neighbour is skipped when its
       triangle equals the voxel's own or an earlier neighbour's, [49]=crossing events,
       [50..65]=evals left per sweep if, in addition, a neighbour is skipped when its triangle has
       not changed since the last sweep in which this voxel looked at the same offset (the 'stamp'
       memo of DESIGN.md; exact because a candidate that lost once can never win later).
       The array has SDFO_NSTATS entries. */
    int64_t *stats;
} sdfo_outputs;

#define SDFO_NSTATS 80
#define IDX(i,j,k) ((int64_t)(i) + (int64_t)ni*((int64_t)(j) + (int64_t)nj*(int64_t)(k)))

static inline v3 vert(const float *x, uint32_t p) { v3 r; memcpy(&r, x + 3*(size_t)p, 12); return r; }

/*
 * Phases A (exact band + crossing counts) and C (sign) for the k-window [k_lo,k_hi) of a grid
 * whose GLOBAL extent is ni x nj x nk; clamps use the global nk (cpu_lib/makelevelset3.cpp:210-212,
 * :222-225).  phi/tri/counts are window-local arrays.  Used directly for slab checks and by the
 * full driver below with the window = whole grid.
 */
static void band_and_counts(const uint32_t *tri, uint64_t ntri, const float *x,
                            const float origin[3], float dx, int ni, int nj, int nk,
                            int k_lo, int k_hi, int band,
                            float *phi, int32_t *ctri, int32_t *cnt, int64_t *stats)
{
    for (uint64_t t = 0; t < ntri; ++t) {
        uint32_t p = tri[3*t], q = tri[3*t+1], r = tri[3*t+2];
        v3 xp = vert(x, p), xq = vert(x, q), xr = vert(x, r);
        /* :206-208 */
        double fip = ((double)xp.v[0]-origin[0])/dx, fjp = ((double)xp.v[1]-origin[1])/dx, fkp = ((double)xp.v[2]-origin[2])/dx;
        double fiq = ((double)xq.v[0]-origin[0])/dx, fjq = ((double)xq.v[1]-origin[1])/dx, fkq = ((double)xq.v[2]-origin[2])/dx;
        double fir = ((double)xr.v[0]-origin[0])/dx, fjr = ((double)xr.v[1]-origin[1])/dx, fkr = ((double)xr.v[2]-origin[2])/dx;
        /* :210-212 */
        int i0 = iclamp((int)dmin3(fip,fiq,fir)-band, 0, ni-1), i1 = iclamp((int)dmax3(fip,fiq,fir)+band+1, 0, ni-1);
        int j0 = iclamp((int)dmin3(fjp,fjq,fjr)-band, 0, nj-1), j1 = iclamp((int)dmax3(fjp,fjq,fjr)+band+1, 0, nj-1);
        int k0 = iclamp((int)dmin3(fkp,fkq,fkr)-band, 0, nk-1), k1 = iclamp((int)dmax3(fkp,fkq,fkr)+band+1, 0, nk-1);
        if (k0 < k_lo) k0 = k_lo;
        if (k1 > k_hi-1) k1 = k_hi-1;
        for (int k = k0; k <= k1; ++k) for (int j = j0; j <= j1; ++j) for (int i = i0; i <= i1; ++i) {
            v3 gx = {{ i*dx+origin[0], j*dx+origin[1], k*dx+origin[2] }};        /* :214 */
            float d = sdfo_point_triangle_distance(gx.v, xp.v, xq.v, xr.v);
            int64_t c = IDX(i, j, k-k_lo);
            if (stats) stats[0]++;
            if (d < phi[c]) { phi[c] = d; ctri[c] = (int32_t)t; }               /* :216-219 */
        }
        /* :222-235 */
        j0 = iclamp((int)ceil(dmin3(fjp,fjq,fjr)), 0, nj-1);
        j1 = iclamp((int)floor(dmax3(fjp,fjq,fjr)), 0, nj-1);
        k0 = iclamp((int)ceil(dmin3(fkp,fkq,fkr)), 0, nk-1);
        k1 = iclamp((int)floor(dmax3(fkp,fkq,fkr)), 0, nk-1);
        if (k0 < k_lo) k0 = k_lo;
        if (k1 > k_hi-1) k1 = k_hi-1;
        for (int k = k0; k <= k1; ++k) for (int j = j0; j <= j1; ++j) {
            double a, b, c;
            if (in_triangle_2d(j, k, fjp, fkp, fjq, fkq, fjr, fkr, &a, &b, &c)) {
                double fi = a*fip + b*fiq + c*fir;
                int ii = (int)ceil(fi);
                if (ii < 0) { ++cnt[IDX(0, j, k-k_lo)]; if (stats) stats[49]++; }
                else if (ii < ni) { ++cnt[IDX(ii, j, k-k_lo)]; if (stats) stats[49]++; }
            }
        }
    }
}

/* :295-303 on a window of nkw planes */
static void apply_sign(int ni, int nj, int nkw, const int32_t *cnt, float *phi)
{
    for (int k = 0; k < nkw; ++k) for (int j = 0; j < nj; ++j) {
        int total = 0;
        for (int i = 0; i < ni; ++i) {
            int64_t c = IDX(i, j, k);
            total += cnt[c];
            if (total % 2 == 1) phi[c] = -phi[c];
        }
    }
}

/* check_neighbour + the pruning counter (:90-102) */
static inline void check_nb(const uint32_t *tri, const float *x, float *phi, int32_t *ctri,
                            v3 gx, int64_t c0, int64_t c1, int64_t *evals)
{
    int32_t t = ctri[c1];
    if (t >= 0) {
        uint32_t p = tri[3*(size_t)t], q = tri[3*(size_t)t+1], r = tri[3*(size_t)t+2];
        float d = sdfo_point_triangle_distance(gx.v, x + 3*(size_t)p, x + 3*(size_t)q, x + 3*(size_t)r);
        if (evals) ++*evals;
        if (d < phi[c0]) { phi[c0] = d; ctri[c0] = t; }
    }
}

/* sweep direction table, cpu_lib/makelevelset3.cpp:245-248 */
static const int SWEEP_DIRS[8][3] = {
    {+1,+1,+1}, {-1,-1,-1}, {+1,+1,-1}, {-1,-1,+1},
    {+1,-1,+1}, {-1,+1,-1}, {+1,-1,-1}, {-1,+1,+1}
};

/* serial sweep, :104-127 (identical visiting order to sweep_range with one thread, :130-151) */
static void sweep_serial(const uint32_t *tri, const float *x, float *phi, int32_t *ctri,
                         const float origin[3], float dx, int ni, int nj, int nk,
                         int di, int dj, int dk, int64_t *stats, int s, uint8_t *stamp)
{
    int i0, i1, j0, j1, k0, k1;
    if (di > 0) { i0 = 1; i1 = ni; } else { i0 = ni-2; i1 = -1; }
    if (dj > 0) { j0 = 1; j1 = nj; } else { j0 = nj-2; j1 = -1; }
    if (dk > 0) { k0 = 1; k1 = nk; } else { k0 = nk-2; k1 = -1; }
    int64_t *ev = stats ? &stats[1+s] : 0;
    /* note: with ni==1 and di<0, i0=-1==i1 and the loop is empty, as in the reference */
    for (int k = k0; k != k1; k += dk) for (int j = j0; j != j1; j += dj) for (int i = i0; i != i1; i += di) {
        v3 gx = {{ i*dx+origin[0], j*dx+origin[1], k*dx+origin[2] }};
        int64_t c = IDX(i, j, k);
        int64_t n[7] = { IDX(i-di,j,k), IDX(i,j-dj,k), IDX(i-di,j-dj,k), IDX(i,j,k-dk),
                         IDX(i-di,j,k-dk), IDX(i,j-dj,k-dk), IDX(i-di,j-dj,k-dk) };   /* :119-125 */
        if (stats) {
            int32_t before = ctri[c], seen[8]; int ns = 0; seen[ns++] = before;
            for (int m = 0; m < 7; ++m) {
                int32_t t = ctri[n[m]];
                if (t < 0) continue;
                int dup = 0; for (int u = 0; u < ns; ++u) if (seen[u] == t) dup = 1;
                if (!dup) {
                    seen[ns++] = t; stats[33+s]++;
                    /* last earlier sweep (1-based) that looked at the same offset m */
                    int last = 0;
                    for (int e = s-1; e >= 0; --e) {
                        const int *dd = SWEEP_DIRS[e % 8];
                        int same = (!(m==0||m==2||m==4||m==6) || dd[0]==di)
                                && (!(m==1||m==2||m==5||m==6) || dd[1]==dj)
                                && (!(m>=3) || dd[2]==dk);
                        if (same) { last = e+1; break; }
                    }
                    if (!(last > 0 && stamp[n[m]] <= last)) stats[50+s]++;
                }
            }
            for (int m = 0; m < 7; ++m) check_nb(tri, x, phi, ctri, gx, c, n[m], ev);
            if (ctri[c] != before) { stats[17+s]++; stamp[c] = (uint8_t)(s+1); }
        } else {
            for (int m = 0; m < 7; ++m) check_nb(tri, x, phi, ctri, gx, c, n[m], 0);
        }
    }
}


/*
 * Full driver, cpu_lib/makelevelset3.cpp:192-304 with num_threads=1.
 * nsweeps: number of direction sweeps to run (16 = the reference's 2 passes x 8; fewer is used
 * by tests that compare after every sweep).  phi_out (required) gets the signed result.
 * Returns 0, or -1 on bad arguments / allocation failure.
 */
int sdfo_make_level_set3(const uint32_t *tri, uint64_t ntri, const float *x, uint64_t nvert,
                         const float origin[3], float dx, int ni, int nj, int nk, int exact_band,
                         int nsweeps, float *phi_out, const sdfo_outputs *out)
{
    (void)nvert;
    if (ni <= 0 || nj <= 0 || nk <= 0 || !phi_out) return -1;
    int64_t V = (int64_t)ni*nj*nk;
    int32_t *ctri = (int32_t*)malloc(sizeof(int32_t)*(size_t)V);
    int32_t *cnt  = (int32_t*)calloc((size_t)V, sizeof(int32_t));
    if (!ctri || !cnt) { free(ctri); free(cnt); return -1; }
    float *phi = phi_out;
    float init = (ni+nj+nk)*dx;                                    /* :197 */
    for (int64_t c = 0; c < V; ++c) { phi[c] = init; ctri[c] = -1; }
    int64_t *stats = out ? out->stats : 0;
    if (stats) memset(stats, 0, sizeof(int64_t)*SDFO_NSTATS);
    uint8_t *stamp = stats ? (uint8_t*)calloc((size_t)V, 1) : 0;

    band_and_counts(tri, ntri, x, origin, dx, ni, nj, nk, 0, nk, exact_band, phi, ctri, cnt, stats);
    if (out && out->phi_band) memcpy(out->phi_band, phi, sizeof(float)*(size_t)V);
    if (out && out->tri_band) memcpy(out->tri_band, ctri, sizeof(int32_t)*(size_t)V);
    if (out && out->counts)   memcpy(out->counts, cnt, sizeof(int32_t)*(size_t)V);

    for (int s = 0; s < nsweeps; ++s) {                            /* :243-292, one thread */
        const int *d = SWEEP_DIRS[s % 8];
        sweep_serial(tri, x, phi, ctri, origin, dx, ni, nj, nk, d[0], d[1], d[2], stats, s < 16 ? s : 15, stamp);
    }
    if (out && out->phi_swept) memcpy(out->phi_swept, phi, sizeof(float)*(size_t)V);
    if (out && out->tri_final) memcpy(out->tri_final, ctri, sizeof(int32_t)*(size_t)V);

    apply_sign(ni, nj, nk, cnt, phi);                              /* :295-303 */
    free(ctri); free(cnt); free(stamp);
    return 0;
}

/*
 * Phases A and C only, on the k-window [k_lo,k_hi) of a (possibly > 2^31-voxel) global grid.
 * Both phases are local in k, so a slab of a grid the reference cannot
 * index can still be checked bit-exactly.  phi_band/tri_band/counts: window-sized outputs.
 */
int sdfo_band_counts_slab(const uint32_t *tri, uint64_t ntri, const float *x,
                          const float origin[3], float dx, int ni, int nj, int nk,
                          int k_lo, int k_hi, int exact_band,
                          float *phi_band, int32_t *tri_band, int32_t *counts)
{
    if (ni <= 0 || nj <= 0 || nk <= 0 || k_lo < 0 || k_hi > nk || k_lo >= k_hi) return -1;
    int64_t V = (int64_t)ni*nj*(k_hi-k_lo);
    float init = (ni+nj+nk)*dx;
    for (int64_t c = 0; c < V; ++c) { phi_band[c] = init; tri_band[c] = -1; counts[c] = 0; }
    band_and_counts(tri, ntri, x, origin, dx, ni, nj, nk, k_lo, k_hi, exact_band, phi_band, tri_band, counts, 0);
    return 0;
}

/* parity of the running crossing count along i, applied to |phi| (window-local) */
void sdfo_apply_sign(int ni, int nj, int nkw, const int32_t *counts, float *phi)
{
    apply_sign(ni, nj, nkw, counts, phi);
}
