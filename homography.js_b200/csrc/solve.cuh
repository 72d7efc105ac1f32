// solve.cuh — the small transform solves, one thread per system, everything in registers.
//
//   affine_from_triangles  : affineMatrixFromTriangles     H.js:1265-1306  (f64 math -> f32[6])
//   inverse_affine         : inverseAffineMatrix           H.js:1345-1365  (f64 math -> f32[6])
//   projective_from_squares: projectiveMatrixFromSquares   H.js:1320-1333  + numeric.js LU / LUsolve
//                            H.js:1664-1749 (8x8, partial pivoting "first strictly larger |a|")
//   transform_limits       : calculateTransformLimits      H.js:1503-1527
//
// Operation order and rounding are the reference's: every multiply, add, subtract and divide is a
// separate IEEE round-to-nearest operation (the __d*_rn intrinsics are never contracted to FMA).
#pragma once
#include "jsnum.cuh"

namespace hg {

__device__ __forceinline__ void affine_from_triangles(const double *s, const double *d, float *out)
{
    const double srcE = s[4], srcF = s[5];
    const double srcA = __dsub_rn(s[0], srcE), srcB = __dsub_rn(s[1], srcF);
    const double srcC = __dsub_rn(s[2], srcE), srcD = __dsub_rn(s[3], srcF);
    const double dstE = d[4], dstF = d[5];
    const double dstA = __dsub_rn(d[0], dstE), dstB = __dsub_rn(d[1], dstF);
    const double dstC = __dsub_rn(d[2], dstE), dstD = __dsub_rn(d[3], dstF);
    const double den = __dsub_rn(__dmul_rn(srcA, srcD), __dmul_rn(srcB, srcC));
    const double nden = -den;
    const double iA = __ddiv_rn(srcD, den);
    const double iB = __ddiv_rn(srcB, nden);
    const double iC = __ddiv_rn(srcC, nden);
    const double iD = __ddiv_rn(srcA, den);
    const double iE = __ddiv_rn(__dsub_rn(__dmul_rn(srcD, srcE), __dmul_rn(srcC, srcF)), nden);
    const double iF = __ddiv_rn(__dsub_rn(__dmul_rn(srcB, srcE), __dmul_rn(srcA, srcF)), den);
    out[0] = __double2float_rn(__dadd_rn(__dmul_rn(dstA, iA), __dmul_rn(dstC, iB)));
    out[1] = __double2float_rn(__dadd_rn(__dmul_rn(dstB, iA), __dmul_rn(dstD, iB)));
    out[2] = __double2float_rn(__dadd_rn(__dmul_rn(dstA, iC), __dmul_rn(dstC, iD)));
    out[3] = __double2float_rn(__dadd_rn(__dmul_rn(dstB, iC), __dmul_rn(dstD, iD)));
    out[4] = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(dstA, iE), __dmul_rn(dstC, iF)), dstE));
    out[5] = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(dstB, iE), __dmul_rn(dstD, iF)), dstF));
}

__device__ __forceinline__ void inverse_affine(const float *m, float *out)
{
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5];
    const double den = __dsub_rn(__dmul_rn(a, d), __dmul_rn(b, c));
    const double nden = -den;
    out[0] = __double2float_rn(__ddiv_rn(d, den));
    out[1] = __double2float_rn(__ddiv_rn(b, nden));
    out[2] = __double2float_rn(__ddiv_rn(c, nden));
    out[3] = __double2float_rn(__ddiv_rn(a, den));
    out[4] = __double2float_rn(__ddiv_rn(__dsub_rn(__dmul_rn(d, e), __dmul_rn(c, f)), nden));
    out[5] = __double2float_rn(__ddiv_rn(__dsub_rn(__dmul_rn(b, e), __dmul_rn(a, f)), den));
}

// 8x8 DLT + LU with partial pivoting + LUsolve.  All loops have compile-time bounds and every
// array index is a compile-time constant after unrolling (row exchanges are predicated swaps), so
// the 64 + 8 doubles live in registers.
__device__ __forceinline__ void projective_from_squares(const double *s, const double *d, double *out)
{
    double A[8][8];
    int P[8];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const double sx = s[2 * p], sy = s[2 * p + 1];
        const double ndx = -d[2 * p], ndy = -d[2 * p + 1];
        A[2 * p][0] = sx; A[2 * p][1] = sy; A[2 * p][2] = 1.0;
        A[2 * p][3] = 0.0; A[2 * p][4] = 0.0; A[2 * p][5] = 0.0;
        A[2 * p][6] = __dmul_rn(ndx, sx); A[2 * p][7] = __dmul_rn(ndx, sy);
        A[2 * p + 1][0] = 0.0; A[2 * p + 1][1] = 0.0; A[2 * p + 1][2] = 0.0;
        A[2 * p + 1][3] = sx; A[2 * p + 1][4] = sy; A[2 * p + 1][5] = 1.0;
        A[2 * p + 1][6] = __dmul_rn(ndy, sx); A[2 * p + 1][7] = __dmul_rn(ndy, sy);
    }
    // Every loop below runs a FIXED 0..7 trip count with an `if` on compile-time indices, so full unrolling is
    // guaranteed and, after it, every array subscript is a literal: A, P and x stay in registers.
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int Pk = k;
        double mx = fabs(A[k][k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j > k) {
                const double a = fabs(A[j][k]);
                if (mx < a) { mx = a; Pk = j; }
            }
        }
        P[k] = Pk;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j > k && Pk == j) {
#pragma unroll
                for (int c = 0; c < 8; ++c) { const double t = A[k][c]; A[k][c] = A[j][c]; A[j][c] = t; }
            }
        }
        const double Akk = A[k][k];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i > k) A[i][k] = __ddiv_rn(A[i][k], Akk);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (i > k && j > k) A[i][j] = __dsub_rn(A[i][j], __dmul_rn(A[i][k], A[k][j]));
        }
    }
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = d[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j > i && P[i] == j) { const double t = x[i]; x[i] = x[j]; x[j] = t; }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < i) x[i] = __dsub_rn(x[i], __dmul_rn(x[j], A[i][j]));
    }
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) {
        const int i = 7 - ii;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j > i) x[i] = __dsub_rn(x[i], __dmul_rn(x[j], A[i][j]));
        x[i] = __ddiv_rn(x[i], A[i][i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = x[i];
}

// applyAffineTransformToPoint / applyProjectiveTransformToPoint for arbitrary double x,y
// (corner evaluation; the per-pixel kernels use specialised forms).
__device__ __forceinline__ void apply_affine_general(const float *m, double x, double y, double &ox, double &oy)
{
    ox = __dadd_rn(__dadd_rn(__dmul_rn((double)m[0], x), __dmul_rn((double)m[2], y)), (double)m[4]);
    oy = __dadd_rn(__dadd_rn(__dmul_rn((double)m[1], x), __dmul_rn((double)m[3], y)), (double)m[5]);
}
__device__ __forceinline__ void apply_projective_general(const double *h, double x, double y, double &ox, double &oy)
{
    const double den = __dadd_rn(__dadd_rn(__dmul_rn(h[6], x), __dmul_rn(h[7], y)), 1.0);
    ox = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(h[0], x), __dmul_rn(h[1], y)), h[2]), den);
    oy = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(h[3], x), __dmul_rn(h[4], y)), h[5]), den);
}

template <int KIND>
__device__ __forceinline__ void transform_limits(const void *matrix, double w, double h, double *out)
{
    double x00, y00, x10, y10, x01, y01, x11, y11;
    if (KIND == 0) {
        const float *m = (const float *)matrix;
        apply_affine_general(m, 0.0, 0.0, x00, y00);
        apply_affine_general(m, 0.0, h, x10, y10);
        apply_affine_general(m, w, 0.0, x01, y01);
        apply_affine_general(m, w, h, x11, y11);
    } else {
        const double *m = (const double *)matrix;
        apply_projective_general(m, 0.0, 0.0, x00, y00);
        apply_projective_general(m, 0.0, h, x10, y10);
        apply_projective_general(m, w, 0.0, x01, y01);
        apply_projective_general(m, w, h, x11, y11);
    }
    const double xo = js_min2(js_min2(js_min2(x00, x10), x01), x11);
    const double yo = js_min2(js_min2(js_min2(y00, y01), y10), y11);
    const double ow = __dsub_rn(js_max2(js_max2(js_max2(x01, x11), x00), x10), xo);
    const double oh = __dsub_rn(js_max2(js_max2(js_max2(y10, y11), y00), y01), yo);
    out[0] = js_round(xo);
    out[1] = js_round(yo);
    out[2] = js_round(ow);
    out[3] = js_round(oh);
}

// ---------------------------------------------------------------- kernels (one thread per system)
// op: 0 = affine solve (in: s[6], d[6] doubles -> f32[6] in out_f)
//     1 = projective solve (in: s[8], d[8] -> f64[8] in out_d)
//     2 = inverse affine (in_f[6] -> out_f[6])
struct SolveArgs {
    const double *src;  // n * (6|8)
    const double *dst;  // n * (6|8)
    const float *in_f;  // n * 6 (op 2)
    float *out_f;       // n * 6
    double *out_d;      // n * 8
    int n;
    int op;
};

__global__ void solve_kernel(SolveArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    if (a.op == 0) {
        double s[6], d[6];
        float o[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) { s[k] = a.src[6 * i + k]; d[k] = a.dst[6 * i + k]; }
        affine_from_triangles(s, d, o);
#pragma unroll
        for (int k = 0; k < 6; ++k) a.out_f[6 * i + k] = o[k];
    } else if (a.op == 1) {
        double s[8], d[8], o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { s[k] = a.src[8 * i + k]; d[k] = a.dst[8 * i + k]; }
        projective_from_squares(s, d, o);
#pragma unroll
        for (int k = 0; k < 8; ++k) a.out_d[8 * i + k] = o[k];
    } else {
        float m[6], o[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) m[k] = a.in_f[6 * i + k];
        inverse_affine(m, o);
#pragma unroll
        for (int k = 0; k < 6; ++k) a.out_f[6 * i + k] = o[k];
    }
}

// one system whose points travel as kernel parameters (no host->device copy of 128 bytes in front of the solve):
// used by the pipelined stream, where a tiny copy between two 8 MB image copies would stall the copy engine
struct SolvePoints {
    double src[8], dst[8];
};
__global__ void solve_points_kernel(const SolvePoints p, int op, float *out_f, double *out_d)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (op == 0) {
        float o[6];
        affine_from_triangles(p.src, p.dst, o);
#pragma unroll
        for (int k = 0; k < 6; ++k) out_f[k] = o[k];
    } else {
        double o[8];
        projective_from_squares(p.src, p.dst, o);
#pragma unroll
        for (int k = 0; k < 8; ++k) out_d[k] = o[k];
    }
}

// limits of one matrix already on the device
__global__ void limits_kernel(int kind, const void *matrix, double w, double h, double *out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (kind == 0) transform_limits<0>(matrix, w, h, out);
    else transform_limits<1>(matrix, w, h, out);
}

}  // namespace hg
