/*
 * hg_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the arithmetic of the image-warp hot path of
 * Eric-Canas/Homography.js (reference file `Homography.js`, cited below as H.js:<line>).
 * It exists so that the CUDA path can be checked bit-for-bit against the reference's
 * semantics on a machine that has no JavaScript engine.  Only tests/, the smoke check
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (homography.js_b200/) never does.
 *
 * Parity pin: reproduces the reference's only pixel golden, test/transformedImage.png,
 * bit-exactly from test/testImgLogoBlack.png + the points of test/nodeTest.js:5-6
 * (see tests/test_oracle_golden.py).  The Delaunay triangulation (third-party
 * `delaunator@5.0.0`, not vendored in the reference) is NOT restated here: triangles
 * are an input (the reference's own injection point is setTriangles, H.js:517), so the
 * piecewise paths are "parity unpinned" at the triangulation boundary only.
 *
 * JavaScript semantics that are reproduced on purpose:
 *   - every Number is an IEEE-754 double, no fused multiply-add (build with
 *     -ffp-contract=off, SSE2 math);
 *   - Math.round = nearest integer, ties toward +infinity;
 *   - ~~x / x<<2 = ToInt32 wrap-around;
 *   - Float32Array stores round to nearest-even float;
 *   - typed-array reads outside [0,len) give undefined (stored as 0 in a
 *     Uint8ClampedArray), writes outside [0,len) are dropped;
 *   - TypedArray.prototype.fill relative-index clamping.
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ JS number helpers */

/* Math.round: integer nearest to x, ties toward +inf; NaN/Inf/huge pass through. */
ORC_API double orc_js_round(double x)
{
    if (!(fabs(x) < 4503599627370496.0)) return x; /* |x| >= 2^52, Inf or NaN */
    double r = floor(x);
    if (x - r >= 0.5) r += 1.0; /* x - floor(x) is exact */
    return r;
}

/* ToInt32 (used by ~~x and by the left operand of <<). */
ORC_API int32_t orc_js_toint32(double x)
{
    if (isnan(x) || isinf(x)) return 0;
    double t = trunc(x);
    double m = fmod(t, 4294967296.0);
    if (m < 0) m += 4294967296.0;
    return (int32_t)(uint32_t)m;
}

/* x << 2 on a JS Number. */
static inline double js_shl2(double x)
{
    return (double)(int32_t)((uint32_t)orc_js_toint32(x) << 2);
}

/* Math.min / Math.max of three values (NaN if any is NaN). */
static inline double js_min2(double a, double b)
{
    if (isnan(a) || isnan(b)) return NAN;
    return a < b ? a : b;
}
static inline double js_max2(double a, double b)
{
    if (isnan(a) || isnan(b)) return NAN;
    return a > b ? a : b;
}
static inline double js_min3(double a, double b, double c) { return js_min2(js_min2(a, b), c); }
static inline double js_max3(double a, double b, double c) { return js_max2(js_max2(a, b), c); }
static inline double js_min4(double a, double b, double c, double d) { return js_min2(js_min3(a, b, c), d); }
static inline double js_max4(double a, double b, double c, double d) { return js_max2(js_max3(a, b, c), d); }

/* ToIntegerOrInfinity + relative-index clamp of TypedArray.prototype.fill. */
static int64_t js_fill_bound(double rel, int64_t len)
{
    if (isnan(rel)) rel = 0.0;
    if (isinf(rel)) return rel < 0 ? 0 : len;
    rel = trunc(rel);
    if (rel < 0) {
        double k = (double)len + rel;
        return k > 0 ? (int64_t)k : 0;
    }
    return rel < (double)len ? (int64_t)rel : len;
}

/* ------------------------------------------------------------------ transform solves */

/* affineMatrixFromTriangles, H.js:1265-1306.  f64 arithmetic, result rounded to f32. */
ORC_API void orc_affine_from_triangles(const double *s, const double *d, float *out)
{
    const double srcE = s[4], srcF = s[5];
    const double srcA = s[0] - srcE, srcB = s[1] - srcF;
    const double srcC = s[2] - srcE, srcD = s[3] - srcF;
    const double dstE = d[4], dstF = d[5];
    const double dstA = d[0] - dstE, dstB = d[1] - dstF;
    const double dstC = d[2] - dstE, dstD = d[3] - dstF;
    const double den = srcA * srcD - srcB * srcC;
    const double iA = srcD / den;
    const double iB = srcB / -den;
    const double iC = srcC / -den;
    const double iD = srcA / den;
    const double iE = (srcD * srcE - srcC * srcF) / -den;
    const double iF = (srcB * srcE - srcA * srcF) / den;
    out[0] = (float)((dstA * iA) + (dstC * iB));
    out[1] = (float)((dstB * iA) + (dstD * iB));
    out[2] = (float)((dstA * iC) + (dstC * iD));
    out[3] = (float)((dstB * iC) + (dstD * iD));
    out[4] = (float)((dstA * iE) + (dstC * iF) + dstE);
    out[5] = (float)((dstB * iE) + (dstD * iF) + dstF);
}

/* inverseAffineMatrix, H.js:1345-1365. */
ORC_API void orc_inverse_affine(const float *m, float *out)
{
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5];
    const double den = a * d - b * c;
    out[0] = (float)(d / den);
    out[1] = (float)(b / -den);
    out[2] = (float)(c / -den);
    out[3] = (float)(a / den);
    out[4] = (float)((d * e - c * f) / -den);
    out[5] = (float)((b * e - a * f) / den);
}

/* numeric.js LU (H.js:1699-1749) + LUsolve (H.js:1664-1697) on the 8x8 DLT system built by
 * projectiveMatrixFromSquares (H.js:1320-1333).  Result: 8 doubles (never rounded to f32). */
ORC_API void orc_projective_from_squares(const double *s, const double *d, double *out)
{
    double A[8][8];
    double *row[8];
    int P[8];
    for (int p = 0; p < 4; ++p) {
        const double sx = s[2 * p], sy = s[2 * p + 1];
        const double dx = d[2 * p], dy = d[2 * p + 1];
        double *r0 = A[2 * p], *r1 = A[2 * p + 1];
        r0[0] = sx; r0[1] = sy; r0[2] = 1; r0[3] = 0; r0[4] = 0; r0[5] = 0;
        r0[6] = -dx * sx; r0[7] = -dx * sy;
        r1[0] = 0; r1[1] = 0; r1[2] = 0; r1[3] = sx; r1[4] = sy; r1[5] = 1;
        r1[6] = -dy * sx; r1[7] = -dy * sy;
    }
    for (int i = 0; i < 8; ++i) row[i] = A[i];
    const int n = 8;
    for (int k = 0; k < n; ++k) {
        int Pk = k;
        double *Ak = row[k];
        double mx = fabs(Ak[k]);
        for (int j = k + 1; j < n; ++j) {
            double a = fabs(row[j][k]);
            if (mx < a) { mx = a; Pk = j; }
        }
        P[k] = Pk;
        if (Pk != k) { row[k] = row[Pk]; row[Pk] = Ak; Ak = row[k]; }
        const double Akk = Ak[k];
        for (int i = k + 1; i < n; ++i) row[i][k] /= Akk;
        for (int i = k + 1; i < n; ++i) {
            double *Ai = row[i];
            for (int j = k + 1; j < n; ++j) Ai[j] -= Ai[k] * Ak[j];
        }
    }
    double x[8];
    for (int i = 0; i < n; ++i) x[i] = d[i];
    for (int i = 0; i < n; ++i) {
        int Pi = P[i];
        if (Pi != i) { double t = x[i]; x[i] = x[Pi]; x[Pi] = t; }
        const double *LUi = row[i];
        for (int j = 0; j < i; ++j) x[i] -= x[j] * LUi[j];
    }
    for (int i = n - 1; i >= 0; --i) {
        const double *LUi = row[i];
        for (int j = i + 1; j < n; ++j) x[i] -= x[j] * LUi[j];
        x[i] /= LUi[i];
    }
    for (int i = 0; i < n; ++i) out[i] = x[i];
}

/* applyAffineTransformToPoint, H.js:1382-1385. */
static inline void apply_affine(const float *m, double x, double y, double *ox, double *oy)
{
    *ox = ((double)m[0] * x) + ((double)m[2] * y) + (double)m[4];
    *oy = ((double)m[1] * x) + ((double)m[3] * y) + (double)m[5];
}
/* applyProjectiveTransformToPoint, H.js:1401-1404. */
static inline void apply_projective(const double *h, double x, double y, double *ox, double *oy)
{
    *ox = (h[0] * x + h[1] * y + h[2]) / (h[6] * x + h[7] * y + 1);
    *oy = (h[3] * x + h[4] * y + h[5]) / (h[6] * x + h[7] * y + 1);
}
ORC_API void orc_apply_affine(const float *m, double x, double y, double *o)
{
    apply_affine(m, x, y, &o[0], &o[1]);
}
ORC_API void orc_apply_projective(const double *h, double x, double y, double *o)
{
    apply_projective(h, x, y, &o[0], &o[1]);
}

/* calculateTransformLimits, H.js:1503-1527.  kind 0 = affine (float[6]), 1 = projective (double[8]).
 * out = [round(xmin), round(ymin), round(xmax-xmin), round(ymax-ymin)] as doubles (may be NaN). */
ORC_API void orc_transform_limits(int kind, const void *matrix, double width, double height, double *out)
{
    double p00[2], p10[2], p01[2], p11[2];
    if (kind == 0) {
        const float *m = (const float *)matrix;
        apply_affine(m, 0, 0, &p00[0], &p00[1]);
        apply_affine(m, 0, height, &p10[0], &p10[1]);
        apply_affine(m, width, 0, &p01[0], &p01[1]);
        apply_affine(m, width, height, &p11[0], &p11[1]);
    } else {
        const double *h = (const double *)matrix;
        apply_projective(h, 0, 0, &p00[0], &p00[1]);
        apply_projective(h, 0, height, &p10[0], &p10[1]);
        apply_projective(h, width, 0, &p01[0], &p01[1]);
        apply_projective(h, width, height, &p11[0], &p11[1]);
    }
    const double xo = js_min4(p00[0], p10[0], p01[0], p11[0]);
    const double yo = js_min4(p00[1], p01[1], p10[1], p11[1]);
    const double ow = js_max4(p01[0], p11[0], p00[0], p10[0]) - xo;
    const double oh = js_max4(p10[1], p11[1], p00[1], p01[1]) - yo;
    out[0] = orc_js_round(xo);
    out[1] = orc_js_round(yo);
    out[2] = orc_js_round(ow);
    out[3] = orc_js_round(oh);
}

/* minmaxXYofArray, H.js:1558-1589: out = [minX, minY, maxX, maxY] (rounded if asked). */
ORC_API void orc_minmax_xy(const double *a, int64_t n, int rounded, double *out)
{
    double maxX = -INFINITY, maxY = -INFINITY, minX = INFINITY, minY = INFINITY;
    for (int64_t i = 0; i < n; ++i) {
        const double e = a[i];
        if ((i % 2) == 0) {
            if (e > maxX) maxX = e;
            if (e < minX) minX = e;
        } else {
            if (e > maxY) maxY = e;
            if (e < minY) minY = e;
        }
    }
    if (rounded) {
        out[0] = orc_js_round(minX); out[1] = orc_js_round(minY);
        out[2] = orc_js_round(maxX); out[3] = orc_js_round(maxY);
    } else {
        out[0] = minX; out[1] = minY; out[2] = maxX; out[3] = maxY;
    }
}

/* ------------------------------------------------------------------ triangle-index map */

typedef struct { double m, b, minY, maxY; } orc_seg;

/* defineTriangleLineEquations, H.js:1141-1151 (triangle = six f32-valued coordinates). */
static void tri_segments(const float *t, orc_seg *s)
{
    const double x0 = t[0], y0 = t[1], x1 = t[2], y1 = t[3], x2 = t[4], y2 = t[5];
    s[0].m = (x1 != x0) ? (y1 - y0) / (x1 - x0) : INFINITY;
    s[0].b = (x1 != x0) ? y0 - x0 * ((y1 - y0) / (x1 - x0)) : x0;
    s[0].minY = js_min2(y1, y0); s[0].maxY = js_max2(y1, y0);
    s[1].m = (x2 != x0) ? (y2 - y0) / (x2 - x0) : INFINITY;
    s[1].b = (x2 != x0) ? y0 - x0 * ((y2 - y0) / (x2 - x0)) : x0;
    s[1].minY = js_min2(y2, y0); s[1].maxY = js_max2(y2, y0);
    s[2].m = (x2 != x1) ? (y2 - y1) / (x2 - x1) : INFINITY;
    s[2].b = (x2 != x1) ? y1 - x1 * ((y2 - y1) / (x2 - x1)) : x1;
    s[2].minY = js_min2(y2, y1); s[2].maxY = js_max2(y2, y1);
}

/* predictXLimits, H.js:1172-1197. */
static void predict_x_limits(const orc_seg *s, double y, double *xmin, double *xmax)
{
    double mn = INFINITY, mx = -INFINITY, x;
    for (int i = 0; i < 3; ++i) {
        if (y >= s[i].minY && y <= s[i].maxY) {
            if (s[i].m == INFINITY) x = s[i].b;
            else if (s[i].m == 0) continue;
            else x = (y - s[i].b) / s[i].m;
            if (x < mn) mn = x;
            if (x > mx) mx = x;
        }
    }
    *xmin = mn; *xmax = mx;
}

/* fillTriangle, H.js:1111-1126.  map is an Int16Array of `len` entries. */
ORC_API void orc_fill_triangle(const float *tri, int32_t idx, double matrix_width, double y_offset,
                               int16_t *map, int64_t len)
{
    const double minY = (double)orc_js_toint32(js_min3(tri[1], tri[3], tri[5]));
    const double maxY = ceil(js_max3(tri[1], tri[3], tri[5]));
    orc_seg seg[3];
    tri_segments(tri, seg);
    const int16_t v = (int16_t)(uint16_t)((uint32_t)idx & 0xFFFFu);
    for (double y = minY; y < maxY; y += 1.0) {
        double xo, xd;
        predict_x_limits(seg, y, &xo, &xd);
        const double start = (y - y_offset) * matrix_width + orc_js_round(xo);
        const double end = (y - y_offset) * matrix_width + orc_js_round(xd);
        const int64_t k0 = js_fill_bound(start, len);
        const int64_t k1 = js_fill_bound(end, len);
        for (int64_t k = k0; k < k1; ++k) map[k] = v;
    }
}

/* _buildTrianglesCorrespondencesMatrix (H.js:817-832) / _buildInverseTrianglesCorrespondencesMatrix
 * (H.js:845-861): fill(-1) then fillTriangle for t = 0..T-1 in order.  pts are the f32 point
 * coordinates (the reference copies them into a Float32Array(6) scratch triangle). */
ORC_API void orc_build_index_map(const float *pts, const uint32_t *tris, int32_t n_tris,
                                 double matrix_width, double y_offset, int16_t *map, int64_t len)
{
    for (int64_t k = 0; k < len; ++k) map[k] = -1;
    for (int32_t t = 0; t < n_tris; ++t) {
        float tri[6];
        for (int v = 0; v < 3; ++v) {
            const uint32_t p = tris[3 * t + v];
            tri[2 * v] = pts[2 * (size_t)p];
            tri[2 * v + 1] = pts[2 * (size_t)p + 1];
        }
        orc_fill_triangle(tri, t, matrix_width, y_offset, map, len);
    }
}

/* _calculatePiecewiseAffineTransformMatrices, H.js:785-804: one forward 2x3 per triangle. */
ORC_API void orc_piecewise_matrices(const float *src_pts, const float *dst_pts, const uint32_t *tris,
                                    int32_t n_tris, float *out /* n_tris*6 */)
{
    for (int32_t t = 0; t < n_tris; ++t) {
        double s[6], d[6];
        for (int v = 0; v < 3; ++v) {
            const uint32_t p = tris[3 * t + v];
            s[2 * v] = src_pts[2 * (size_t)p]; s[2 * v + 1] = src_pts[2 * (size_t)p + 1];
            d[2 * v] = dst_pts[2 * (size_t)p]; d[2 * v + 1] = dst_pts[2 * (size_t)p + 1];
        }
        orc_affine_from_triangles(s, d, out + 6 * (size_t)t);
    }
}

/* ------------------------------------------------------------------ warp loops */

static inline void copy_px_read_guard(uint8_t *out, int64_t out_idx, int64_t out_len,
                                      const uint8_t *img, double src_idx, int64_t img_len)
{
    /* out[out_idx+c] = image[src_idx+c]; OOB read -> undefined -> 0; OOB write -> dropped */
    if (out_idx >= 0 && out_idx + 3 < out_len && src_idx >= 0 && src_idx + 3 < (double)img_len) {
        memcpy(out + out_idx, img + (int64_t)src_idx, 4); /* common case: 4 in-range byte copies */
        return;
    }
    for (int c = 0; c < 4; ++c) {
        const int64_t oi = out_idx + c;
        if (oi < 0 || oi >= out_len) continue;
        const double si = src_idx + c;
        uint8_t v = 0;
        if (si >= 0 && si < (double)img_len && si == trunc(si)) v = img[(int64_t)si];
        out[oi] = v;
    }
}

/* _inverseGeometricWarp, H.js:987-1013.  kind 0 affine (float[6]) / 1 projective (double[8]).
 * `inv` is the already-solved inverse (dst->src) matrix.  out must hold oW*oH*4 bytes; it is
 * zero-filled here (new Uint8ClampedArray).  threads<=1 -> the reference's single thread. */
ORC_API void orc_warp_inverse_geometric(int kind, const uint8_t *img, int32_t W, int32_t H, const void *inv,
                                        int32_t xOff, int32_t yOff, int32_t oW, int32_t oH, uint8_t *out,
                                        int threads)
{
    const double srcRow = (double)(int32_t)((uint32_t)W << 2);
    const double dstRow = (double)(int32_t)((uint32_t)oW << 2);
    const int64_t out_len = (int64_t)(dstRow * (double)oH);
    const int64_t img_len = (int64_t)W * H * 4;
    if (out_len <= 0) return;
    memset(out, 0, (size_t)out_len);
    const float *ma = (const float *)inv;
    const double *mp = (const double *)inv;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads > 1 ? threads : 1) if (threads > 1)
#endif
    for (int32_t yy = 0; yy < oH; ++yy) {
        const double y = (double)yOff + (double)yy;
        for (int32_t xx = 0; xx < oW; ++xx) {
            const double x = (double)xOff + (double)xx;
            double sx, sy;
            if (kind == 0) apply_affine(ma, x, y, &sx, &sy);
            else apply_projective(mp, x, y, &sx, &sy);
            if (sx >= 0 && sx < (double)W && sy >= 0 && sy < (double)H) {
                const int64_t idx = (int64_t)(((y - yOff) * dstRow) + js_shl2(x - xOff));
                const double sidx = (orc_js_round(sy) * srcRow) + js_shl2(orc_js_round(sx));
                copy_px_read_guard(out, idx, out_len, img, sidx, img_len);
            }
        }
    }
}

/* EXTENSION (not in the reference, which only has Math.round sampling — SURVEY Q12): bilinear sampling for the inverse
 * affine / projective loop.  Definition used as the oracle ("parity unpinned": the reference has nothing to pin it to):
 * same window test as H.js:1001 on the unrounded (sx,sy); x0 = floor(sx), fx = sx - x0, x1 = min(x0+1, W-1) (edge
 * replicate), likewise y; channel = p00(1-fx)(1-fy) + p10 fx(1-fy) + p01(1-fx)fy + p11 fx fy in double; stored like a
 * Uint8ClampedArray store (round half to even, clamp). */
ORC_API void orc_warp_inverse_geometric_bilinear(int kind, const uint8_t *img, int32_t W, int32_t H, const void *inv,
                                                 int32_t xOff, int32_t yOff, int32_t oW, int32_t oH, uint8_t *out,
                                                 int threads)
{
    const int64_t out_len = (int64_t)oW * oH * 4;
    if (out_len <= 0) return;
    memset(out, 0, (size_t)out_len);
    const float *ma = (const float *)inv;
    const double *mp = (const double *)inv;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads > 1 ? threads : 1) if (threads > 1)
#endif
    for (int32_t yy = 0; yy < oH; ++yy) {
        const double y = (double)yOff + (double)yy;
        for (int32_t xx = 0; xx < oW; ++xx) {
            const double x = (double)xOff + (double)xx;
            double sx, sy;
            if (kind == 0) apply_affine(ma, x, y, &sx, &sy);
            else apply_projective(mp, x, y, &sx, &sy);
            if (sx >= 0 && sx < (double)W && sy >= 0 && sy < (double)H) {
                const double fx0 = floor(sx), fy0 = floor(sy);
                const double fx = sx - fx0, fy = sy - fy0;
                const int64_t x0 = (int64_t)fx0, y0 = (int64_t)fy0;
                const int64_t x1 = x0 + 1 < W ? x0 + 1 : W - 1, y1 = y0 + 1 < H ? y0 + 1 : H - 1;
                const uint8_t *p00 = img + 4 * (y0 * W + x0), *p10 = img + 4 * (y0 * W + x1);
                const uint8_t *p01 = img + 4 * (y1 * W + x0), *p11 = img + 4 * (y1 * W + x1);
                uint8_t *o = out + 4 * ((int64_t)yy * oW + xx);
                for (int c = 0; c < 4; ++c) {
                    const double v = p00[c] * (1 - fx) * (1 - fy) + p10[c] * fx * (1 - fy) + p01[c] * (1 - fx) * fy +
                                     p11[c] * fx * fy;
                    double r = nearbyint(v); /* round half to even (default rounding mode) */
                    if (r < 0) r = 0;
                    if (r > 255) r = 255;
                    o[c] = (uint8_t)r;
                }
            }
        }
    }
}

/* _geometricWarp, H.js:911-932 (forward scatter; source raster order, last writer wins). */
ORC_API void orc_warp_forward_geometric(int kind, const uint8_t *img, int32_t W, int32_t H, const void *fwd,
                                        int32_t xOff, int32_t yOff, int32_t oW, int32_t oH, uint8_t *out)
{
    const double srcRow = (double)(int32_t)((uint32_t)W << 2);
    const double dstRow = (double)(int32_t)((uint32_t)oW << 2);
    const int64_t out_len = (int64_t)(dstRow * (double)oH);
    const int64_t img_len = (int64_t)W * H * 4;
    if (out_len <= 0) return;
    memset(out, 0, (size_t)out_len);
    const float *ma = (const float *)fwd;
    const double *mp = (const double *)fwd;
    for (int32_t yi = 0; yi < H; ++yi) {
        for (int32_t xi = 0; xi < W; ++xi) {
            const double x = xi, y = yi;
            const double idx = (y * srcRow) + js_shl2(x);
            double nx, ny;
            if (kind == 0) apply_affine(ma, x, y, &nx, &ny);
            else apply_projective(mp, x, y, &nx, &ny);
            nx = orc_js_round(nx - xOff);
            ny = orc_js_round(ny - yOff);
            const double nidx = (ny * dstRow) + js_shl2(nx);
            if (!(nidx >= 0 && nidx < (double)out_len)) continue; /* NaN or OOB: every write dropped */
            copy_px_read_guard(out, (int64_t)nidx, out_len, img, idx, img_len);
        }
    }
}

/* _inversePiecewiseAffineWarp pixel loop, H.js:1042-1056.  map = inverse index map (Int16),
 * inv = T inverse 2x3 float matrices (H.js:1036-1038 applied to the forward ones). */
ORC_API void orc_warp_inverse_piecewise(const uint8_t *img, int32_t W, int32_t H, const int16_t *map,
                                        int64_t map_len, const float *inv, int32_t n_tris, int32_t xOff,
                                        int32_t yOff, int32_t oW, int32_t oH, int32_t minSrcX, int32_t minSrcY,
                                        uint8_t *out, int threads)
{
    const double srcRow = (double)(int32_t)((uint32_t)W << 2);
    const double dstRow = (double)(int32_t)((uint32_t)oW << 2);
    const int64_t out_len = (int64_t)(dstRow * (double)oH);
    const int64_t img_len = (int64_t)W * H * 4;
    if (out_len <= 0) return;
    memset(out, 0, (size_t)out_len);
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads > 1 ? threads : 1) if (threads > 1)
#endif
    for (int32_t yy = 0; yy < oH; ++yy) {
        const double y = (double)yOff + (double)yy;
        for (int32_t xx = 0; xx < oW; ++xx) {
            const double x = (double)xOff + (double)xx;
            const int64_t mi = (int64_t)((y - yOff) * (double)oW + (x - xOff));
            if (mi < 0 || mi >= map_len) continue; /* undefined >= 0 is false */
            const int32_t t = map[mi];
            if (t >= 0) {
                if (t >= n_tris) continue; /* reference would throw on undefined matrix; unreachable for valid maps */
                double sx, sy;
                apply_affine(inv + 6 * (size_t)t, x, y, &sx, &sy);
                if (sx >= minSrcX && sx < (double)W + minSrcX && sy >= minSrcY && sy < (double)H + minSrcY) {
                    sx = orc_js_round(sx); sy = orc_js_round(sy);
                    const double sidx = (sy * srcRow) + js_shl2(sx);
                    const int64_t didx = (int64_t)(((y - yOff) * dstRow) + js_shl2(x - xOff));
                    copy_px_read_guard(out, didx, out_len, img, sidx, img_len);
                }
            }
        }
    }
}

/* _piecewiseAffineWarp, H.js:948-972 (forward scatter over the src-points bbox). */
ORC_API void orc_warp_forward_piecewise(const uint8_t *img, int32_t W, int32_t H, const int16_t *map,
                                        int64_t map_len, const float *fwd, int32_t n_tris, int32_t xOff,
                                        int32_t yOff, int32_t oW, int32_t oH, int32_t minSrcX, int32_t minSrcY,
                                        int32_t maxSrcX, int32_t maxSrcY, uint8_t *out)
{
    const double srcRow = (double)(int32_t)((uint32_t)W << 2);
    const double dstRow = (double)(int32_t)((uint32_t)oW << 2);
    const int64_t out_len = (int64_t)(dstRow * (double)oH);
    const int64_t img_len = (int64_t)W * H * 4;
    const double mw = (double)maxSrcX - (double)minSrcX;
    if (out_len <= 0) return;
    memset(out, 0, (size_t)out_len);
    for (int32_t yi = minSrcY; yi < maxSrcY; ++yi) {
        for (int32_t xi = minSrcX; xi < maxSrcX; ++xi) {
            const double x = xi, y = yi;
            const double mid = (y - minSrcY) * mw + (x - minSrcX);
            if (!(mid >= 0 && mid < (double)map_len)) continue; /* undefined > -1 is false */
            const int32_t t = map[(int64_t)mid];
            if (t > -1) {
                if (t >= n_tris) continue;
                const double idx = (y * srcRow) + js_shl2(x);
                double nx, ny;
                apply_affine(fwd + 6 * (size_t)t, x, y, &nx, &ny);
                nx = orc_js_round(nx - xOff);
                ny = orc_js_round(ny - yOff);
                const double nidx = (ny * dstRow) + js_shl2(nx);
                if (!(nidx >= 0 && nidx < (double)out_len)) continue;
                copy_px_read_guard(out, (int64_t)nidx, out_len, img, idx, img_len);
            }
        }
    }
}

ORC_API int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
