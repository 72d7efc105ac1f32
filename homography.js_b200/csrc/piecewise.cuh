// piecewise.cuh — piecewise-affine path, GENERAL form (exact for every input the reference accepts).
//
//   pw_setup_kernel   one thread per triangle: forward 2x3 (affineMatrixFromTriangles, H.js:1265, via
//                     _calculatePiecewiseAffineTransformMatrices H.js:785), its inverse
//                     (inverseAffineMatrix H.js:1345, as H.js:1036-1038), the three edge equations of the
//                     triangle that is rasterised (defineTriangleLineEquations H.js:1141) and its row range
//                     [~~minY, ceil(maxY)) (H.js:1113-1115)
//   pw_fill_kernel    fillTriangle / predictXLimits (H.js:1111-1126, 1172-1197) + TypedArray.fill index
//                     semantics.  The reference fills triangles sequentially so the LAST (highest) triangle
//                     index wins; here every (triangle,row) span is written with atomicMax on a 32-bit map,
//                     an order-independent reduction with the same result
//   pw_warp_inverse_kernel   pixel loop of _inversePiecewiseAffineWarp (H.js:1042-1056)
//   map32_to_int16_kernel    Int16Array store semantics of the map (H.js:820/848: ids wrap mod 2^16)
//
// The fused, map-free path lives in piecewise_fused.cuh and reuses pw_setup_kernel / predict_x_limits from here.
// The map-based kernels below are (a) the fallback for frames the bins of the fused path cannot represent (more
// than 8 overlapping spans in one 64-pixel block, >= 2^17 triangles), (b) the forward map of _piecewiseAffineWarp
// and (c) hg_build_index_map, i.e. the Int16 map itself for callers / tests that want it.
#pragma once
#include "solve.cuh"

namespace hg {

struct TriRec {
    float fwd[6];
    float inv[6];
    double m[3], b[3], lo[3], hi[3];  // edge equations: slope, intercept (or x for vertical), minY, maxY
    double maxY;                      // ceil(max y)
    int y0;                           // ~~min y
    int pad;
};

struct PwSetupArgs {
    const float *src_pts;   // mesh source points
    const float *dst_pts;   // destiny points (may be nullptr when only a map is wanted)
    const float *map_pts;   // the points whose triangles are rasterised (dst for the inverse map)
    const uint32_t *tris;
    TriRec *rec;
    float *fwd_out;         // optional dense copies (T*6)
    float *inv_out;
    float *invd_out;        // optional: inverse matrices as 32-byte records (T*8 floats per frame), for the fused kernel
    int2 *yr_out;           // optional: the rows [y0, y_end) each triangle's fill loop visits (T per frame), for the band binning pass
    int n_tris;
    size_t dst_stride;      // frame stride (floats) for batched calls, indexed by blockIdx.y
    size_t rec_stride;
};

__global__ void pw_setup_kernel(PwSetupArgs a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_tris) return;
    const size_t f = blockIdx.y;
    const uint32_t i0 = a.tris[3 * t], i1 = a.tris[3 * t + 1], i2 = a.tris[3 * t + 2];
    TriRec r;
    if (a.dst_pts) {
        const float *dp = a.dst_pts + f * a.dst_stride;
        double s[6], d[6];
        s[0] = a.src_pts[2 * i0]; s[1] = a.src_pts[2 * i0 + 1];
        s[2] = a.src_pts[2 * i1]; s[3] = a.src_pts[2 * i1 + 1];
        s[4] = a.src_pts[2 * i2]; s[5] = a.src_pts[2 * i2 + 1];
        d[0] = dp[2 * i0]; d[1] = dp[2 * i0 + 1];
        d[2] = dp[2 * i1]; d[3] = dp[2 * i1 + 1];
        d[4] = dp[2 * i2]; d[5] = dp[2 * i2 + 1];
        affine_from_triangles(s, d, r.fwd);
        inverse_affine(r.fwd, r.inv);
    } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) { r.fwd[k] = 0.f; r.inv[k] = 0.f; }
    }
    {
        const float *mp = a.map_pts + (a.map_pts == a.dst_pts ? f * a.dst_stride : 0);
        const double x0 = mp[2 * i0], y0 = mp[2 * i0 + 1];
        const double x1 = mp[2 * i1], y1 = mp[2 * i1 + 1];
        const double x2 = mp[2 * i2], y2 = mp[2 * i2 + 1];
        const double inf = __longlong_as_double(0x7ff0000000000000LL);
        // p0->p1, p0->p2, p1->p2  (H.js:1145-1150)
        const double xa[3] = {x0, x0, x1}, ya[3] = {y0, y0, y1};
        const double xb[3] = {x1, x2, x2}, yb[3] = {y1, y2, y2};
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            if (xb[e] != xa[e]) {
                const double slope = __ddiv_rn(__dsub_rn(yb[e], ya[e]), __dsub_rn(xb[e], xa[e]));
                r.m[e] = slope;
                r.b[e] = __dsub_rn(ya[e], __dmul_rn(xa[e], slope));
            } else {
                r.m[e] = inf;
                r.b[e] = xa[e];
            }
            r.lo[e] = js_min2(yb[e], ya[e]);
            r.hi[e] = js_max2(yb[e], ya[e]);
        }
        r.y0 = js_toint32(js_min2(js_min2(y0, y1), y2));
        r.maxY = ceil(js_max2(js_max2(y0, y1), y2));
        r.pad = 0;
    }
    a.rec[f * a.rec_stride + t] = r;
    if (a.fwd_out) {
#pragma unroll
        for (int k = 0; k < 6; ++k) a.fwd_out[(f * a.n_tris + t) * 6 + k] = r.fwd[k];
    }
    if (a.inv_out) {
#pragma unroll
        for (int k = 0; k < 6; ++k) a.inv_out[(f * a.n_tris + t) * 6 + k] = r.inv[k];
    }
    // H.js:1114-1116: for (y = y0; y < maxY; y++) with maxY = Math.ceil(max y), an integer below 2^21 (points are validated)
    if (a.yr_out) a.yr_out[f * a.n_tris + t] = make_int2(r.y0, (r.maxY > (double)r.y0) ? (int)r.maxY : r.y0);
    if (a.invd_out) {
        float4 *o = reinterpret_cast<float4 *>(a.invd_out + (f * a.n_tris + t) * 8);
        o[0] = make_float4(r.inv[0], r.inv[1], r.inv[2], r.inv[3]);
        o[1] = make_float4(r.inv[4], r.inv[5], 0.f, 0.f);
    }
}

// predictXLimits (H.js:1172) for one row of one triangle
__device__ __forceinline__ void predict_x_limits(const TriRec &r, double y, double &xmin, double &xmax)
{
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double mn = inf, mx = -inf;
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        if (y >= r.lo[e] && y <= r.hi[e]) {
            double x;
            if (r.m[e] == inf) x = r.b[e];
            else if (r.m[e] == 0.0) continue;
            else x = __ddiv_rn(__dsub_rn(y, r.b[e]), r.m[e]);
            if (x < mn) mn = x;
            if (x > mx) mx = x;
        }
    }
    xmin = mn;
    xmax = mx;
}

struct PwFillArgs {
    const TriRec *rec;
    int *map32;        // initialised to -1
    long long map_len;
    double map_width;
    double y_offset;
    int n_tris;
    int row_split;     // gridDim.y: block (t, s) takes rows r with (r / 8) % row_split == s
};

// block = 32 x 8: threadIdx.y picks the row inside a group of 8, lanes stride along the span
__global__ void __launch_bounds__(256) pw_fill_kernel(PwFillArgs a)
{
    const int t = blockIdx.x;
    const TriRec &r = a.rec[t];
    const double maxY = r.maxY;
    for (long long grp = blockIdx.y;; grp += a.row_split) {
        const double y = (double)r.y0 + (double)(grp * 8 + threadIdx.y);
        const double ybase = (double)r.y0 + (double)(grp * 8);
        if (!(ybase < maxY)) break;  // also ends on NaN
        if (!(y < maxY)) continue;
        double xo, xd;
        predict_x_limits(r, y, xo, xd);
        const double rowbase = __dmul_rn(__dsub_rn(y, a.y_offset), a.map_width);
        const long long k0 = js_fill_bound(__dadd_rn(rowbase, js_round(xo)), a.map_len);
        const long long k1 = js_fill_bound(__dadd_rn(rowbase, js_round(xd)), a.map_len);
        for (long long k = k0 + threadIdx.x; k < k1; k += 32) atomicMax(a.map32 + k, t);
    }
}

__global__ void map32_to_int16_kernel(const int *map32, short *out, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int t = map32[i];
        out[i] = (t < 0) ? (short)-1 : (short)(unsigned short)(t & 0xFFFF);
    }
}

// A9 for piecewise frames (H.js:706-710): xOff = round(min x), yOff = round(min y), oW = round(max x) - xOff,
// oH = round(max y) - yOff — the difference of ROUNDED extrema (Q13).  One warp per frame; `>`/`<` skip NaN exactly
// like minmaxXYofArray (H.js:1558).  out[f] = {xOff, yOff, oW, oH} as doubles (JS Numbers: may be +-Inf).
__global__ void __launch_bounds__(128) pw_extent_kernel(const float *dst_pts, int n_pts, int n_frames, double *out)
{
    const int f = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (f >= n_frames) return;
    const int lane = threadIdx.x & 31;
    const float *p = dst_pts + (size_t)f * 2 * n_pts;
    const float inf = __int_as_float(0x7f800000);
    float mnx = inf, mny = inf, mxx = -inf, mxy = -inf;
    for (int i = lane; i < n_pts; i += 32) {
        const float x = p[2 * i], y = p[2 * i + 1];
        if (x > mxx) mxx = x;
        if (x < mnx) mnx = x;
        if (y > mxy) mxy = y;
        if (y < mny) mny = y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if (lane == 0) {
        const double x0 = js_round((double)mnx), y0 = js_round((double)mny);
        out[4 * f + 0] = x0;
        out[4 * f + 1] = y0;
        out[4 * f + 2] = __dsub_rn(js_round((double)mxx), x0);
        out[4 * f + 3] = __dsub_rn(js_round((double)mxy), y0);
    }
}

struct PwWarpArgs {
    const uint32_t *src;
    uint32_t *out;
    const int *map32;
    const TriRec *rec;
    int W, H, xOff, yOff, oW, oH;
    int minSrcX, minSrcY;
    int n_tris;
};

// pixel loop of _inversePiecewiseAffineWarp (H.js:1042-1056); 4 pixels / thread, one 128-bit store
__global__ void __launch_bounds__(256) pw_warp_inverse_kernel(const PwWarpArgs a)
{
    const long long npix = (long long)a.oW * a.oH;
    const long long nquad = (npix + 3) >> 2;
    const long long npx_src = (long long)a.W * a.H;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nquad; q += stride) {
        const long long p0 = q << 2;
        int yy = (int)(p0 / a.oW);
        int xx = (int)(p0 - (long long)yy * a.oW);
        int tq[4];
        if (p0 + 3 < npix) {
            const int4 tv = *reinterpret_cast<const int4 *>(a.map32 + p0);
            tq[0] = tv.x; tq[1] = tv.y; tq[2] = tv.z; tq[3] = tv.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) tq[k] = (p0 + k < npix) ? a.map32[p0 + k] : -1;
        }
        uint32_t px[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t v = 0u;
            // Int16Array semantics: the stored id is t mod 2^16 reinterpreted as int16
            const int t = (tq[k] < 0) ? -1 : (int)(short)(unsigned short)(tq[k] & 0xFFFF);
            if (p0 + k < npix && t >= 0 && t < a.n_tris) {
                const float *m = a.rec[t].inv;
                const double x = (double)(a.xOff + xx), y = (double)(a.yOff + yy);
                const double sx = affine_coord_exact((double)m[0], x, __dmul_rn((double)m[2], y), (double)m[4]);
                const double sy = affine_coord_exact((double)m[1], x, __dmul_rn((double)m[3], y), (double)m[5]);
                const FloorHalf fx = floor_half_exact(sx);
                const FloorHalf fy = floor_half_exact(sy);
                // minSrcX <= sx < W + minSrcX  and  minSrcY <= sy < H + minSrcY   (H.js:1047)
                if (fx.ok && fy.ok && (unsigned)(fx.ipart - a.minSrcX) < (unsigned)a.W &&
                    (unsigned)(fy.ipart - a.minSrcY) < (unsigned)a.H) {
                    const long long flat = (long long)round_half_up(fy) * a.W + round_half_up(fx);
                    if (flat >= 0 && flat < npx_src) v = __ldg(a.src + flat);
                }
            }
            px[k] = v;
            if (++xx == a.oW) { xx = 0; ++yy; }
        }
        if (p0 + 3 < npix) {
            *reinterpret_cast<uint4 *>(a.out + p0) = make_uint4(px[0], px[1], px[2], px[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (p0 + k < npix) a.out[p0 + k] = px[k];
        }
    }
}

}  // namespace hg
