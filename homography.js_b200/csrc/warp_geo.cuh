// warp_geo.cuh — K1/K2: the pixel loop of _inverseGeometricWarp (H.js:997-1011) as an HBM-bound
// gather kernel.
//
//   for every output pixel (x,y) in [xOff,xOff+oW) x [yOff,yOff+oH):
//       (sx,sy) = T^-1(x,y)                      affine: H.js:1382, projective: H.js:1401
//       if 0 <= sx < W and 0 <= sy < H:          H.js:1001 (test on the UNROUNDED coordinate)
//           out[x,y] = src_flat[round(sy)*W + round(sx)]   H.js:1005 (flat index, Math.round;
//                                                 index past the end reads `undefined` -> 0)
//
// Layout: RGBA8 pixels are handled as one 32-bit word.  The output is addressed FLAT
// (pixel p = yy*oW + xx); one thread produces 4 consecutive pixels and issues ONE 128-bit store,
// so a warp writes 512 contiguous bytes regardless of oW.  Source reads are 32-bit read-only
// gathers (adjacent output pixels map to adjacent source pixels, so a warp's 128 loads fall in a
// handful of 128-byte lines and the image stays L2-resident across the frame).
//
// Arithmetic: bit-exact with the reference's unfused doubles.
//   affine      2 DFMA + 2 DADD + 2 DADD.RD per pixel (products float x int are exact, see jsnum.cuh)
//   projective  numerators / denominator evaluated exactly as the reference does (DMUL + DADD);
//               the two IEEE divides are replaced by ONE Newton reciprocal + 2 DMUL whose relative
//               error is < 2^-48; every decision of the loop flips only at multiples of 0.5, so
//               the quotient is trusted unless it lies within 2^-24 of such a multiple (probability
//               ~2.4e-7 per coordinate), in which case the pixel is redone with __ddiv_rn.
#pragma once
#include "jsnum.cuh"

namespace hg {

struct GeoFrame {
    const uint32_t *src;
    uint32_t *out;
    int W, H;
    int xOff, yOff, oW, oH;
};

struct GeoParams {
    GeoFrame one;          // used when many == nullptr
    const GeoFrame *many;  // device array, indexed by blockIdx.y
    const void *mats_dev;  // device matrices (float[6] | double[8] per frame) or nullptr
    double mat_val[8];     // matrix by value when mats_dev == nullptr (floats widened to double)
};

__device__ __forceinline__ uint32_t fetch_src(const uint32_t *__restrict__ src, int W, long long npx_src,
                                              const FloorHalf &fx, const FloorHalf &fy)
{
    // srcIdx = round(sy)*W + round(sx) in pixel units; >= W*H reads `undefined` -> 0 (H.js:1005-1007)
    const long long flat = (long long)round_half_up(fy) * W + round_half_up(fx);
    return (flat < npx_src) ? __ldg(src + flat) : 0u;
}

// rcp.approx.ftz.f64 (MUFU.RCP64H, ~20 good bits) + two Newton steps: relative error < 2^-48 for
// normal d.  NOT correctly rounded: callers must use the near-boundary filter.
__device__ __forceinline__ double rcp_newton(double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = __fma_rn(-d, r, 1.0);
    r = __fma_rn(r, e, r);
    e = __fma_rn(-d, r, 1.0);
    r = __fma_rn(r, e, r);
    return r;
}

#define HG_NEAR_DELTA 256u  // 2^8 * 2^-32 = 2^-24 absolute

template <int KIND>
__global__ void __launch_bounds__(256) warp_inverse_geo_kernel(const GeoParams P)
{
    const GeoFrame F = P.many ? P.many[blockIdx.y] : P.one;
    const long long npix = (long long)F.oW * F.oH;
    const long long nquad = (npix + 3) >> 2;
    const long long npx_src = (long long)F.W * F.H;

    double m[8];
    if (P.mats_dev) {
        if (KIND == 0) {
            const float *mf = (const float *)P.mats_dev + 6 * (size_t)blockIdx.y;
#pragma unroll
            for (int k = 0; k < 6; ++k) m[k] = (double)__ldg(mf + k);
        } else {
            const double *md = (const double *)P.mats_dev + 8 * (size_t)blockIdx.y;
#pragma unroll
            for (int k = 0; k < 8; ++k) m[k] = __ldg(md + k);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = P.mat_val[k];
    }

    const uint32_t *__restrict__ src = F.src;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nquad; q += stride) {
        const long long p0 = q << 2;
        int yy = (int)(p0 / F.oW);
        int xx = (int)(p0 - (long long)yy * F.oW);
        uint32_t px[4];
        // row-constant terms (recomputed when the quad crosses a row end)
        double y = (double)(F.yOff + yy);
        double r0, r1, r2;
        if (KIND == 0) {
            r0 = __dmul_rn(m[2], y);
            r1 = __dmul_rn(m[3], y);
            r2 = 0.0;
        } else {
            r0 = __dmul_rn(m[1], y);
            r1 = __dmul_rn(m[4], y);
            r2 = __dmul_rn(m[7], y);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t v = 0u;
            if (p0 + k < npix) {
                const double x = (double)(F.xOff + xx);
                if (KIND == 0) {
                    const double sx = affine_coord_exact(m[0], x, r0, m[4]);
                    const double sy = affine_coord_exact(m[1], x, r1, m[5]);
                    const FloorHalf fx = floor_half_exact(sx);
                    const FloorHalf fy = floor_half_exact(sy);
                    if (fx.ok && fy.ok && (unsigned)fx.ipart < (unsigned)F.W && (unsigned)fy.ipart < (unsigned)F.H)
                        v = fetch_src(src, F.W, npx_src, fx, fy);
                } else {
                    const double nx = __dadd_rn(__dadd_rn(__dmul_rn(m[0], x), r0), m[2]);
                    const double ny = __dadd_rn(__dadd_rn(__dmul_rn(m[3], x), r1), m[5]);
                    const double dn = __dadd_rn(__dadd_rn(__dmul_rn(m[6], x), r2), 1.0);
                    // fast quotients
                    const double rc = rcp_newton(dn);
                    bool nearx, neary;
                    FloorHalf fx = floor_half_approx(__dmul_rn(nx, rc), HG_NEAR_DELTA, nearx);
                    FloorHalf fy = floor_half_approx(__dmul_rn(ny, rc), HG_NEAR_DELTA, neary);
                    // |dn| outside [2^-500, 2^500] (or 0 / Inf / NaN): the reciprocal is not trusted
                    const unsigned de = ((unsigned)__double2hiint(dn) >> 20) & 0x7FFu;
                    const bool d_bad = (de - 523u) > 1000u;
                    if (d_bad || (fx.ok && nearx) || (fy.ok && neary)) {
                        fx = floor_half_exact(__ddiv_rn(nx, dn));  // exact path == the reference
                        fy = floor_half_exact(__ddiv_rn(ny, dn));
                    }
                    if (fx.ok && fy.ok && (unsigned)fx.ipart < (unsigned)F.W && (unsigned)fy.ipart < (unsigned)F.H)
                        v = fetch_src(src, F.W, npx_src, fx, fy);
                }
                if (++xx == F.oW) {  // next pixel starts a new output row
                    xx = 0;
                    ++yy;
                    y = (double)(F.yOff + yy);
                    if (KIND == 0) {
                        r0 = __dmul_rn(m[2], y);
                        r1 = __dmul_rn(m[3], y);
                    } else {
                        r0 = __dmul_rn(m[1], y);
                        r1 = __dmul_rn(m[4], y);
                        r2 = __dmul_rn(m[7], y);
                    }
                }
            }
            px[k] = v;
        }
        if (p0 + 3 < npix) {
            *reinterpret_cast<uint4 *>(F.out + p0) = make_uint4(px[0], px[1], px[2], px[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (p0 + k < npix) F.out[p0 + k] = px[k];
        }
    }
}

}  // namespace hg
