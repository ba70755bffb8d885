// sdfb_math.cuh -- bit-exact device arithmetic for make_level_set3.
//
// Every function here must produce the same bits as the reference's x86-64 build
// (/root/reference/cpu_lib/makelevelset3.cpp compiled -O3 without FMA): each multiply and add is
// rounded separately (explicit __f*_rn / __d*_rn intrinsics are never contracted by nvcc, and the
// translation units are also built with -fmad=false), division and square root are IEEE
// round-to-nearest (-prec-div=true -prec-sqrt=true -ftz=false, the nvcc defaults), and the
// std::min/std::max NaN-ordering of the reference is kept.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sdfb {

struct F3 { float x, y, z; };

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }

__device__ __forceinline__ F3 sub3(F3 a, F3 b) { return F3{fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)}; }
// common/vec.h:377-383: ((a0*b0)+(a1*b1))+(a2*b2)
__device__ __forceinline__ float dot3(F3 a, F3 b) { return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z)); }
// common/vec.h:216-223
__device__ __forceinline__ float mag2_3(F3 a) { return dot3(a, a); }
// common/vec.h:240-255
__device__ __forceinline__ float dist3(F3 a, F3 b) { return __fsqrt_rn(mag2_3(sub3(a, b))); }
// std::min(a,b) = (b<a)?b:a ; std::max(a,b) = (a<b)?b:a   (common/util.h:22-23)
__device__ __forceinline__ float  min_std(float a, float b)   { return (b < a) ? b : a; }
__device__ __forceinline__ float  max_std(float a, float b)   { return (a < b) ? b : a; }
__device__ __forceinline__ double min_std(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double max_std(double a, double b) { return (a < b) ? b : a; }

// cpu_lib/makelevelset3.cpp:21-34.  The reference narrows a double quotient of two floats,
// (float)((double)dot/(double)m2); because double carries 53 >= 2*24+2 bits that double rounding is
// innocuous and equals the correctly rounded float quotient (checked exhaustively-at-random on the
// CPU, tests/test_oracle.py::test_float_div_equals_narrowed_double_div), so one fp32 IEEE divide is
// used here instead of an fp64 one.
// Returns the SQUARED distance; the callers take one square root of the smaller of two candidates, which
// is bit-identical to the reference's min(sqrt(a), sqrt(b)): sqrt is monotone and correctly rounded, and the
// std::min tie/NaN ordering is the same on the squares (a NaN operand wins or loses identically).
__device__ __forceinline__ float seg_distance2(F3 x0, F3 x1, F3 x2)
{
    F3 e = sub3(x2, x1);
    float m2 = mag2_3(e);
    float s12 = __fdiv_rn(dot3(sub3(x2, x0), e), m2);
    if (s12 < 0.f) s12 = 0.f; else if (s12 > 1.f) s12 = 1.f;
    float om = fsub(1.f, s12);
    F3 p = F3{ fadd(fmul(s12, x1.x), fmul(om, x2.x)),
               fadd(fmul(s12, x1.y), fmul(om, x2.y)),
               fadd(fmul(s12, x1.z), fmul(om, x2.z)) };
    return mag2_3(sub3(x0, p));
}
__device__ __forceinline__ float seg_distance(F3 x0, F3 x1, F3 x2) { return __fsqrt_rn(seg_distance2(x0, x1, x2)); }

// cpu_lib/makelevelset3.cpp:49-70
__device__ __forceinline__ float point_triangle_distance(F3 x0, F3 x1, F3 x2, F3 x3)
{
    F3 x13 = sub3(x1, x3), x23 = sub3(x2, x3), x03 = sub3(x0, x3);
    float m13 = mag2_3(x13), m23 = mag2_3(x23), d = dot3(x13, x23);
    float invdet = __fdiv_rn(1.f, max_std(fsub(fmul(m13, m23), fmul(d, d)), 1e-30f));
    float a = dot3(x13, x03), b = dot3(x23, x03);
    float w23 = fmul(invdet, fsub(fmul(m23, a), fmul(d, b)));
    float w31 = fmul(invdet, fsub(fmul(m13, b), fmul(d, a)));
    float w12 = fsub(fsub(1.f, w23), w31);
    float d2;
    if (w23 >= 0.f && w31 >= 0.f && w12 >= 0.f) {
        F3 p = F3{ fadd(fadd(fmul(w23, x1.x), fmul(w31, x2.x)), fmul(w12, x3.x)),
                   fadd(fadd(fmul(w23, x1.y), fmul(w31, x2.y)), fmul(w12, x3.y)),
                   fadd(fadd(fmul(w23, x1.z), fmul(w31, x2.z)), fmul(w12, x3.z)) };
        d2 = mag2_3(sub3(x0, p));
    } else if (w23 > 0.f) {
        d2 = min_std(seg_distance2(x0, x1, x2), seg_distance2(x0, x1, x3));
    } else if (w31 > 0.f) {
        d2 = min_std(seg_distance2(x0, x1, x2), seg_distance2(x0, x2, x3));
    } else {
        d2 = min_std(seg_distance2(x0, x1, x3), seg_distance2(x0, x2, x3));
    }
    return __fsqrt_rn(d2);
}

// World position of lattice point c along one axis: float(c)*dx + origin  (cpu_lib/makelevelset3.cpp:214)
__device__ __forceinline__ float lattice(int c, float dx, float o) { return fadd(fmul(__int2float_rn(c), dx), o); }

// (int) of a double as the reference's x86-64 build performs it (cvttsd2si): truncation toward zero,
// 0x80000000 for NaN / out-of-range.  In-range values are all that well-formed inputs produce.
__device__ __forceinline__ int d2i_trunc(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0)) return (int)0x80000000;
    return __double2int_rz(v);
}
__device__ __forceinline__ int iclamp(int a, int lo, int hi) { return a < lo ? lo : (a > hi ? hi : a); }
// wrap-around int subtract/add, as two's-complement hardware does (avoids signed-overflow UB)
__device__ __forceinline__ int wrap_add(int a, int b) { return (int)((unsigned)a + (unsigned)b); }

// cpu_lib/makelevelset3.cpp:155-165
__device__ __forceinline__ int orientation(double x1, double y1, double x2, double y2, double &twice_signed_area)
{
    twice_signed_area = __dsub_rn(__dmul_rn(y1, x2), __dmul_rn(x1, y2));
    if (twice_signed_area > 0) return 1;
    else if (twice_signed_area < 0) return -1;
    else if (y2 > y1) return 1;
    else if (y2 < y1) return -1;
    else if (x1 > x2) return 1;
    else if (x1 < x2) return -1;
    else return 0;
}

// cpu_lib/makelevelset3.cpp:169-187
__device__ __forceinline__ bool point_in_triangle_2d(double x0, double y0, double x1, double y1,
                                                     double x2, double y2, double x3, double y3,
                                                     double &a, double &b, double &c)
{
    x1 = __dsub_rn(x1, x0); x2 = __dsub_rn(x2, x0); x3 = __dsub_rn(x3, x0);
    y1 = __dsub_rn(y1, y0); y2 = __dsub_rn(y2, y0); y3 = __dsub_rn(y3, y0);
    int signa = orientation(x2, y2, x3, y3, a);
    if (signa == 0) return false;
    int signb = orientation(x3, y3, x1, y1, b);
    if (signb != signa) return false;
    int signc = orientation(x1, y1, x2, y2, c);
    if (signc != signa) return false;
    double sum = __dadd_rn(__dadd_rn(a, b), c);
    a = __ddiv_rn(a, sum); b = __ddiv_rn(b, sum); c = __ddiv_rn(c, sum);
    return true;
}

// ---- cell packing --------------------------------------------------------------------------
// cell = (float bits of |phi|) << 32 | (stamp << 27) | (closest_tri & 0x07ffffff)
//   closest_tri == -1 is stored as the 27-bit all-ones pattern; stamp = index (1..16) of the last
//   sweep that changed the cell's triangle, 0 for the exact band; the initial cell is
//   (bits(init_phi) << 32) | 0xffffffff.  Distances are >= +0 so the unsigned 64-bit order of cells
//   with stamp 0 is the lexicographic order of (phi, closest_tri): atomicMin resolves the exact band
//   to the reference's serial result (lowest distance, then lowest triangle index).
constexpr uint32_t TRI_MASK = 0x07ffffffu;
constexpr uint32_t TRI_NONE = 0x07ffffffu;
__device__ __forceinline__ uint64_t pack_cell(float phi, uint32_t lo) { return ((uint64_t)__float_as_uint(phi) << 32) | lo; }
__device__ __forceinline__ float    cell_phi(uint64_t c) { return __uint_as_float((uint32_t)(c >> 32)); }
__device__ __forceinline__ uint32_t cell_lo(uint64_t c) { return (uint32_t)c; }
__device__ __forceinline__ uint32_t lo_tri(uint32_t lo) { return lo & TRI_MASK; }
__device__ __forceinline__ uint32_t lo_stamp(uint32_t lo) { return lo >> 27; }
__device__ __forceinline__ int32_t  lo_tri_signed(uint32_t lo) { uint32_t t = lo & TRI_MASK; return t == TRI_NONE ? -1 : (int32_t)t; }

// Per-triangle record: the three vertices pre-gathered as 3 x float4 (48 B, 16-byte aligned) so a
// candidate evaluation is three LDG.128 instead of an index load plus three scattered 12-byte loads.
// The three spare lanes carry per-triangle invariants of point_triangle_distance, computed once with the
// same operations in the same order (so the bits are those of the reference): p.w = m13 = |x1-x3|^2,
// q.w = m23 = |x2-x3|^2, r.w = invdet = 1/max(m13*m23 - d*d, 1e-30f)  (cpu_lib/makelevelset3.cpp:52-54).
struct __align__(16) TriRec { float4 p, q, r; };

__device__ __forceinline__ TriRec make_tri_rec(F3 x1, F3 x2, F3 x3)
{
    F3 x13 = sub3(x1, x3), x23 = sub3(x2, x3);
    float m13 = mag2_3(x13), m23 = mag2_3(x23), d = dot3(x13, x23);
    float invdet = __fdiv_rn(1.f, max_std(fsub(fmul(m13, m23), fmul(d, d)), 1e-30f));
    TriRec t;
    t.p = make_float4(x1.x, x1.y, x1.z, m13);
    t.q = make_float4(x2.x, x2.y, x2.z, m23);
    t.r = make_float4(x3.x, x3.y, x3.z, invdet);
    return t;
}

// point_triangle_distance with the record's invariants (bit-identical to the full function).
// The three "clamp to two edges" cases of the reference (:63-68) always measure two of the segments
// (x1,x2), (x1,x3), (x2,x3) in that orientation and take min(first, second); instead of branching three ways
// (lanes of a warp disagree on the case all the time) the endpoints are SELECTED and the two segment
// distances are computed once, uniformly.  Same operations on the same values per lane => same bits.
__device__ __forceinline__ F3 sel3(bool c, F3 a, F3 b) { return F3{c ? a.x : b.x, c ? a.y : b.y, c ? a.z : b.z}; }
__device__ __forceinline__ float ptd_rec(F3 x0, float4 P, float4 Q, float4 R)
{
    const F3 x1{P.x, P.y, P.z}, x2{Q.x, Q.y, Q.z}, x3{R.x, R.y, R.z};
    const float m13 = P.w, m23 = Q.w, invdet = R.w;
    F3 x13 = sub3(x1, x3), x23 = sub3(x2, x3), x03 = sub3(x0, x3);
    float d = dot3(x13, x23);
    float a = dot3(x13, x03), b = dot3(x23, x03);
    float w23 = fmul(invdet, fsub(fmul(m23, a), fmul(d, b)));
    float w31 = fmul(invdet, fsub(fmul(m13, b), fmul(d, a)));
    float w12 = fsub(fsub(1.f, w23), w31);
    float d2;
    if (w23 >= 0.f && w31 >= 0.f && w12 >= 0.f) {
        F3 p = F3{ fadd(fadd(fmul(w23, x1.x), fmul(w31, x2.x)), fmul(w12, x3.x)),
                   fadd(fadd(fmul(w23, x1.y), fmul(w31, x2.y)), fmul(w12, x3.y)),
                   fadd(fadd(fmul(w23, x1.z), fmul(w31, x2.z)), fmul(w12, x3.z)) };
        d2 = mag2_3(sub3(x0, p));
    } else {
        const bool c1 = w23 > 0.f, c12 = c1 || (w31 > 0.f);
        // first segment: (x1,x2) in cases 1 and 2, (x1,x3) in case 3; second: (x1,x3) in case 1, else (x2,x3)
        const F3 fb = sel3(c12, x2, x3);
        const F3 sa = sel3(c1, x1, x2);
        d2 = min_std(seg_distance2(x0, x1, fb), seg_distance2(x0, sa, x3));
    }
    return __fsqrt_rn(d2);
}
__device__ __forceinline__ float ptd_rec(F3 gx, const TriRec &t) { return ptd_rec(gx, t.p, t.q, t.r); }

}  // namespace sdfb
