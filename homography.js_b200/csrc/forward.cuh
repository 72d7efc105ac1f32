// forward.cuh — the forward-scatter loops, made deterministic.
//
//   _geometricWarp        H.js:911-932   loop over SOURCE pixels (x,y) in raster order:
//                                        (nx,ny) = round(T(x,y) - offset); out[ny*4oW + (nx<<2) ..+3] = src[y*4W + (x<<2) ..+3]
//   _piecewiseAffineWarp  H.js:948-972   same over the integer bounding box of the source points, through the
//                                        forward index map and the per-triangle forward matrix
//
// The reference runs sequentially, so when several source pixels land on one output pixel the LAST one in loop
// order wins, and writes whose flat index falls outside [0, len) are dropped while in-range-but-wrapped ones land
// (Q3).  Here: pass 1 scatters the loop-order KEY of every source element with atomicMax into a 32-bit "winner"
// plane (an order-independent reduction with the same result), pass 2 gathers the winning source pixel for every
// output pixel and writes the whole output once with 128-bit stores (untouched pixels become transparent).
#pragma once
#include "piecewise.cuh"

namespace hg {

// Target pixel index of a forward-mapped point, exactly as the reference computes it:
//   newX = Math.round(tx - xOff); newY = Math.round(ty - yOff); newIdx = newY*(oW<<2) + (newX<<2)
// Returns -1 when the write is dropped (NaN / outside [0, len)).
__device__ __forceinline__ long long forward_target(double tx, double ty, int xOff, int yOff, int oW, long long npix_out)
{
    const double dx = __dsub_rn(tx, (double)xOff), dy = __dsub_rn(ty, (double)yOff);
    const FloorHalf fx = floor_half_exact(dx), fy = floor_half_exact(dy);
    if (fx.ok && fy.ok) {  // |coordinates| < 2^19: plain integer arithmetic is the same thing
        const long long p = (long long)round_half_up(fy) * oW + round_half_up(fx);
        return (p >= 0 && p < npix_out) ? p : -1;
    }
    // general JS semantics (huge / NaN / Inf coordinates): ToInt32 wrap of newX<<2, double arithmetic for the sum
    const double rx = js_round(dx), ry = js_round(dy);
    const double dst_row = (double)(int)((unsigned)oW << 2);
    const double sh = (double)(int)((unsigned)js_toint32(rx) << 2);
    const double idx = __dadd_rn(__dmul_rn(ry, dst_row), sh);
    if (!(idx >= 0.0 && idx < 4.0 * (double)npix_out)) return -1;
    if (idx != trunc(idx)) return -1;
    return (long long)idx >> 2;  // idx is a multiple of 4 whenever it is an in-range integer
}

struct FwdArgs {
    const uint32_t *src;
    uint32_t *out;
    int *winner;             // oW*oH ints, initialised to -1
    const int *map32;        // piecewise only
    const TriRec *rec;       // piecewise only
    long long map_len;
    double mat[8];           // geometric only (affine floats widened / projective doubles)
    int kind;                // geometric: 0 affine, 1 projective
    int W, H, xOff, yOff, oW, oH;
    int minX, minY, domW, domH;  // loop domain: x in [minX, minX+domW), y in [minY, minY+domH)
    int n_tris;
};

template <bool PIECEWISE>
__global__ void __launch_bounds__(256) forward_scatter_kernel(const FwdArgs a)
{
    const long long n = (long long)a.domW * a.domH;
    const long long npix_out = (long long)a.oW * a.oH;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long key = (long long)blockIdx.x * blockDim.x + threadIdx.x; key < n; key += stride) {
        const int yy = (int)(key / a.domW);
        const int xx = (int)(key - (long long)yy * a.domW);
        const double x = (double)(a.minX + xx), y = (double)(a.minY + yy);
        double tx, ty;
        if (PIECEWISE) {
            if (key >= a.map_len) continue;  // read past the map: undefined > -1 is false
            const int raw = a.map32[key];
            const int t = (raw < 0) ? -1 : (int)(short)(unsigned short)(raw & 0xFFFF);  // Int16Array semantics
            if (t < 0 || t >= a.n_tris) continue;
            apply_affine_general(a.rec[t].fwd, x, y, tx, ty);
        } else if (a.kind == 0) {
            float m[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) m[k] = (float)a.mat[k];
            apply_affine_general(m, x, y, tx, ty);
        } else {
            apply_projective_general(a.mat, x, y, tx, ty);
        }
        const long long p = forward_target(tx, ty, a.xOff, a.yOff, a.oW, npix_out);
        if (p >= 0) atomicMax(a.winner + p, (int)key);
    }
}

// pass 2: every output pixel takes the source pixel of its winning key (or stays transparent)
__global__ void __launch_bounds__(256) forward_gather_kernel(const FwdArgs a)
{
    const long long npix = (long long)a.oW * a.oH;
    const long long nquad = (npix + 3) >> 2;
    const long long npx_src = (long long)a.W * a.H;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nquad; q += stride) {
        const long long p0 = q << 2;
        uint32_t px[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t v = 0u;
            if (p0 + k < npix) {
                const int key = a.winner[p0 + k];
                if (key >= 0) {
                    const int yy = key / a.domW;
                    const int xx = key - yy * a.domW;
                    // idx = y*(W<<2) + (x<<2): flat, so x >= W runs into the next row; outside the image -> 0
                    const long long flat = (long long)(a.minY + yy) * a.W + (a.minX + xx);
                    if (flat >= 0 && flat < npx_src) v = __ldg(a.src + flat);
                }
            }
            px[k] = v;
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
