// jsnum.cuh — JavaScript Number semantics on sm_100a, bit for bit.
//
// The reference (Homography.js) computes every coordinate in IEEE-754 double WITHOUT fused
// multiply-add, rounds with Math.round (ties toward +inf) and indexes typed arrays with the
// result.  Everything here is written with explicit round-to-nearest intrinsics (__dmul_rn,
// __dadd_rn, __ddiv_rn never contract), and FMA appears only where it is provably identical
// (both products are exact, see affine_coord_exact()).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hg {

// 1.5 * 2^20.  For a double v with -2^19 <= v < 2^19, t = v + kMagic lies in [2^20, 2^21) where
// ulp = 2^-32, so the 52 mantissa bits of t are the fixed-point value (v + 2^19) * 2^32:
//   high word bits 0..19 = floor(v) + 2^19,  low word = frac(v) * 2^32.
#define HG_MAGIC 1572864.0
#define HG_MAGIC_EXP 0x413  // biased exponent of [2^20, 2^21), sign bit clear

struct FloorHalf {
    int ipart;      // floor(v)                      (valid iff ok)
    unsigned frac;  // frac(v) * 2^32, truncated     (valid iff ok)
    bool ok;        // v finite and -2^19 <= v < 2^19
};

// EXACT floor / fraction of v.  The add rounds toward -inf: every integer and every half-integer
// is on the 2^-32 grid, so rounding down to the grid can change neither floor(v) nor the truth of
// frac(v) >= 0.5.  One FP64-pipe instruction replaces floor(), the bounds compares and Math.round.
__device__ __forceinline__ FloorHalf floor_half_exact(double v)
{
    const double t = __dadd_rd(v, HG_MAGIC);
    const int hi = __double2hiint(t);
    FloorHalf r;
    r.frac = (unsigned)__double2loint(t);
    r.ok = (hi >> 20) == HG_MAGIC_EXP;
    r.ipart = (hi & 0xFFFFF) - (1 << 19);
    return r;
}

// Math.round(v) for a decomposed v: floor(v) + (frac(v) >= 0.5)
__device__ __forceinline__ int round_half_up(const FloorHalf &f) { return f.ipart + (int)(f.frac >> 31); }

// General Math.round (any double, incl. NaN / Inf / huge) — cold paths only.
__device__ __forceinline__ double js_round(double x)
{
    if (!(fabs(x) < 4503599627370496.0)) return x;
    double r = floor(x);
    if (__dsub_rn(x, r) >= 0.5) r = __dadd_rn(r, 1.0);
    return r;
}

// ToInt32 (~~x, and the operand conversion of x << 2).
__device__ __forceinline__ int js_toint32(double x)
{
    if (isnan(x) || isinf(x)) return 0;
    double t = trunc(x);
    double m = fmod(t, 4294967296.0);
    if (m < 0) m += 4294967296.0;
    return (int)(unsigned)m;
}

// Math.min / Math.max: NaN if any operand is NaN.
__device__ __forceinline__ double js_min2(double a, double b)
{
    if (isnan(a) || isnan(b)) return __longlong_as_double(0x7ff8000000000000LL);
    return a < b ? a : b;
}
__device__ __forceinline__ double js_max2(double a, double b)
{
    if (isnan(a) || isnan(b)) return __longlong_as_double(0x7ff8000000000000LL);
    return a > b ? a : b;
}

// TypedArray.prototype.fill relative index -> absolute index in [0, len].
__device__ __forceinline__ long long js_fill_bound(double rel, long long len)
{
    if (isnan(rel)) rel = 0.0;
    if (isinf(rel)) return rel < 0 ? 0 : len;
    rel = trunc(rel);
    if (rel < 0) {
        double k = __dadd_rn((double)len, rel);
        return k > 0 ? (long long)k : 0;
    }
    return rel < (double)len ? (long long)rel : len;
}

// applyAffineTransformToPoint (H.js:1382) for float coefficients and integer-valued |x|,|y| < 2^28:
//   (m0*x + m2*y) + m4.  m*x is a 24-bit by <=28-bit product, exact in double, so
//   fma(m0, x, m2*y) == RN(RN(m0*x) + RN(m2*y)) — the same bits as the unfused reference.
__device__ __forceinline__ double affine_coord_exact(double m_x, double x, double m_y_times_y, double m_c)
{
    return __dadd_rn(__fma_rn(m_x, x, m_y_times_y), m_c);
}

}  // namespace hg
