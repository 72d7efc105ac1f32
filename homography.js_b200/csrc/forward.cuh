// forward.cuh — the forward-scatter loops, made deterministic.
//
//   _geometricWarp        H.js:911-932   loop over SOURCE pixels (x,y) in raster order:
//                                        (nx,ny) = round(T(x,y) - offset); out[ny*4oW + (nx<<2) ..+3] = src[y*4W + (x<<2) ..+3]
//   _piecewiseAffineWarp  H.js:948-972   same over the integer bounding box of the source points, through the
//                                        forward index map and the per-triangle forward matrix
//
// The reference runs sequentially, so when several source pixels land on one output pixel the LAST one in loop
// order wins, and writes whose flat index falls outside [0, len) are dropped while in-range-but-wrapped ones land
// (Q3).  Here:
//   general     pass 1 scatters the loop-order KEY of every source element with a 32-bit atomic max into a "winner"
//               plane (an order-independent reduction with the same result; RED, no return value), pass 2 gathers the
//               winning source pixel for every output pixel, writes the whole output once with 128-bit stores
//               (untouched pixels become transparent) and RESETS the plane entry it read — so the plane is clean for
//               the next frame without a memset pass, and a small ring of planes stays resident in the 126 MB L2
//               while a batch streams through it (DRAM sees the 4 B read + 4 B write per pixel only).
//   lattice     affine matrices whose linear part is a signed permutation with exact 0 / +-1 entries (translations,
//               mirrors, quarter turns: what warp() actually dispatches here — the forward loop is only chosen when
//               the output has the size of the input, H.js:426) map the pixel lattice onto itself one to one:
//               round((+-x) + e - xOff) = +-x + round(e - xOff) exactly.  No collisions, no ordering question: one
//               gather pass with the integer inverse, 8 B per pixel, no atomics (forward_lattice_kernel; the host
//               proves the preconditions per frame, see forward_lattice_plan in hgwarp.cu).
// Both kernels take a batch: blockIdx.y = frame.
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
    int *winner;             // oW*oH ints, all -1 between frames
    const int *map32;        // piecewise only
    const TriRec *rec;       // piecewise only
    long long map_len;
    double mat[8];           // geometric only (affine floats widened / projective doubles)
    int kind;                // geometric: 0 affine, 1 projective
    int W, H, xOff, yOff, oW, oH;
    int minX, minY, domW, domH;  // loop domain: x in [minX, minX+domW), y in [minY, minY+domH)
    int n_tris;
    // lattice plan (forward_lattice_kernel): source pixel of output (X, Y) is
    //   x = ixx * (X - rx) + ixy * (Y - ry),  y = iyx * (X - rx) + iyy * (Y - ry)     (integer inverse of the permutation)
    int lattice;             // 1: this frame takes the lattice kernel
    int ixx, ixy, iyx, iyy, rx, ry;
    int shift_copy;          // lattice frame that is a pure translation with 16-byte aligned rows on both sides
};

struct FwdParams {
    FwdArgs one;             // used when many == nullptr
    const FwdArgs *many;     // device array, indexed by blockIdx.y
};

template <bool PIECEWISE>
__global__ void __launch_bounds__(256) forward_scatter_kernel(const FwdParams P)
{
    const FwdArgs &a = P.many ? P.many[blockIdx.y] : P.one;
    if (a.lattice) return;
    const int n = a.domW * a.domH;  // < 2^31 (checked on the host)
    const long long npix_out = (long long)a.oW * a.oH;
    const int stride = (int)(gridDim.x * blockDim.x);
    float mf[6];
    if (!PIECEWISE) {
#pragma unroll
        for (int k = 0; k < 6; ++k) mf[k] = (float)a.mat[k];
    }
    // key - stride may wrap past 2^31 only after the loop condition has failed: the unsigned compare ends the loop
    for (unsigned key = blockIdx.x * blockDim.x + threadIdx.x; key < (unsigned)n; key += (unsigned)stride) {
        const int yy = (int)(key / (unsigned)a.domW);
        const int xx = (int)(key - (unsigned)yy * (unsigned)a.domW);
        const double x = (double)(a.minX + xx), y = (double)(a.minY + yy);
        double tx, ty;
        if (PIECEWISE) {
            if ((long long)key >= a.map_len) continue;  // read past the map: undefined > -1 is false
            const int raw = __ldg(a.map32 + key);
            const int t = (raw < 0) ? -1 : (int)(short)(unsigned short)(raw & 0xFFFF);  // Int16Array semantics
            if (t < 0 || t >= a.n_tris) continue;
            apply_affine_general(a.rec[t].fwd, x, y, tx, ty);
        } else if (a.kind == 0) {
            apply_affine_general(mf, x, y, tx, ty);
        } else {
            apply_projective_general(a.mat, x, y, tx, ty);
        }
        const long long p = forward_target(tx, ty, a.xOff, a.yOff, a.oW, npix_out);
        if (p >= 0) atomicMax(a.winner + p, (int)key);
    }
}

// pass 2: every output pixel takes the source pixel of its winning key (or stays transparent) and hands the plane
// entry back as -1
__global__ void __launch_bounds__(256) forward_gather_kernel(const FwdParams P)
{
    const FwdArgs &a = P.many ? P.many[blockIdx.y] : P.one;
    if (a.lattice) return;
    const long long npix = (long long)a.oW * a.oH;
    const long long nquad = (npix + 3) >> 2;
    const long long npx_src = (long long)a.W * a.H;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nquad; q += stride) {
        const long long p0 = q << 2;
        int key[4];
        if (p0 + 3 < npix) {
            int4 *wp = reinterpret_cast<int4 *>(a.winner + p0);
            const int4 kv = *wp;
            key[0] = kv.x; key[1] = kv.y; key[2] = kv.z; key[3] = kv.w;
            if ((kv.x & kv.y & kv.z & kv.w) != -1) *wp = make_int4(-1, -1, -1, -1);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                key[k] = -1;
                if (p0 + k < npix) {
                    key[k] = a.winner[p0 + k];
                    if (key[k] != -1) a.winner[p0 + k] = -1;
                }
            }
        }
        uint32_t px[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t v = 0u;
            if (key[k] >= 0) {
                const int yy = key[k] / a.domW;
                const int xx = key[k] - yy * a.domW;
                // idx = y*(W<<2) + (x<<2): flat, so x >= W runs into the next row; outside the image -> 0
                const long long flat = (long long)(a.minY + yy) * a.W + (a.minX + xx);
                if (flat >= 0 && flat < npx_src) v = __ldg(a.src + flat);
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

// lattice frames: out(X, Y) = src(x, y) with the integer inverse of the plan, transparent where (x, y) leaves the
// loop domain [0, W) x [0, H).  Flat quads (a quad may run over a row end: (X, Y) per pixel); a thread takes TWO quads
// per step, a grid stride apart, and issues the eight gathers before the two 128-bit stores.
__device__ __forceinline__ void lattice_quad_load(const FwdArgs &a, const uint32_t *__restrict__ src, int p0, int npix, uint32_t (&px)[4])
{
    int Y = p0 / a.oW;
    int X = p0 - Y * a.oW;
    if (a.shift_copy) {
        // a translation between images whose rows are whole 16-byte groups: the quad is four consecutive source pixels
        // x0 .. x0+3 of one row — the aligned group that holds x0 and (unless x0 is aligned itself) the next one, i.e. one
        // or two 128-bit loads instead of four 32-bit ones through the same sectors
        const int y = Y - a.ry, x0 = X - a.rx;
        if ((unsigned)y < (unsigned)a.H && x0 >= 0 && x0 + 7 < a.W) {
            const uint4 *g = reinterpret_cast<const uint4 *>(src + (size_t)y * a.W + (x0 & ~3));
            const uint4 A = __ldg(g);
            const int off = x0 & 3;   // the same for every quad of the frame
            if (off == 0) {
                px[0] = A.x; px[1] = A.y; px[2] = A.z; px[3] = A.w;
            } else {
                const uint4 B = __ldg(g + 1);
                if (off == 1) { px[0] = A.y; px[1] = A.z; px[2] = A.w; px[3] = B.x; }
                else if (off == 2) { px[0] = A.z; px[1] = A.w; px[2] = B.x; px[3] = B.y; }
                else { px[0] = A.w; px[1] = B.x; px[2] = B.y; px[3] = B.z; }
            }
            return;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int u = X - a.rx, v = Y - a.ry;
        const int x = a.ixx * u + a.ixy * v, y = a.iyx * u + a.iyy * v;
        uint32_t w = 0u;
        if ((unsigned)x < (unsigned)a.W && (unsigned)y < (unsigned)a.H && p0 + k < npix) w = __ldg(src + (size_t)y * a.W + x);
        px[k] = w;
        if (++X == a.oW) { X = 0; ++Y; }
    }
}
__device__ __forceinline__ void lattice_quad_store(uint32_t *out, int p0, int npix, const uint32_t (&px)[4])
{
    if (p0 + 3 < npix) {
        *reinterpret_cast<uint4 *>(out + p0) = make_uint4(px[0], px[1], px[2], px[3]);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (p0 + k < npix) out[p0 + k] = px[k];
    }
}

__global__ void __launch_bounds__(256) forward_lattice_kernel(const FwdParams P)
{
    const FwdArgs &a = P.many ? P.many[blockIdx.y] : P.one;
    if (!a.lattice) return;
    const int npix = a.oW * a.oH;  // < 2^31 (checked on the host)
    const int nquad = (npix + 3) >> 2;
    const int stride = (int)(gridDim.x * blockDim.x);
    const uint32_t *__restrict__ src = a.src;
    for (int q = (int)(blockIdx.x * blockDim.x + threadIdx.x); q < nquad; q += 2 * stride) {
        uint32_t pa[4], pb[4];
        const int q2 = q + stride;
        lattice_quad_load(a, src, q << 2, npix, pa);
        if (q2 < nquad) lattice_quad_load(a, src, q2 << 2, npix, pb);
        lattice_quad_store(a.out, q << 2, npix, pa);
        if (q2 < nquad) lattice_quad_store(a.out, q2 << 2, npix, pb);
    }
}

// fills a buffer of ints with -1 (a fresh winner plane)
__global__ void fill_minus_one_kernel(int *p, long long n)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = -1;
}

}  // namespace hg
