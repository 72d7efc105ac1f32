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

// pass 1.  A WARP takes 128 consecutive source pixels of one row of the loop domain per step, lane l the pixels l, l + 32,
// l + 64, l + 96: the 32 atomics of one instruction then land on 32 neighbouring targets for maps near the identity (four
// 32-byte sectors, i.e. four requests to the L2 atomic units) instead of 32 targets four pixels apart (sixteen).  One
// division per warp step, the row terms m2*y / m3*y once per step; a piecewise lane fetches a triangle's forward matrix only
// when the triangle changes.  The frame's constants are copied to registers first: through the descriptor reference every
// field would be re-read after each atomic (the compiler must assume the atomic wrote it) — 8 % of the first version's
// instructions were such loads.  Every product and sum is the reference's own (H.js:1382-1385 / 1401-1404): unfused, in
// its order.  MODE: 0 affine, 1 projective (host-dispatched: a batch has one kind), 2 piecewise.
template <int MODE>
__global__ void __launch_bounds__(256) forward_scatter_kernel(const FwdParams P)
{
    const FwdArgs &g = P.many ? P.many[blockIdx.y] : P.one;
    if (g.lattice) return;
    const int domW = g.domW, domH = g.domH, minX = g.minX, minY = g.minY, oW = g.oW, xOff = g.xOff, yOff = g.yOff, n_tris = g.n_tris;
    int *const winner = g.winner;
    const int *const map32 = g.map32;
    const TriRec *const rec = g.rec;
    const long long map_len = g.map_len;
    const long long npix_out = (long long)oW * g.oH;
    double m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = MODE == 2 ? 0.0 : ((MODE == 0 && k < 6) ? (double)(float)g.mat[k] : g.mat[k]);
    const int nseg_row = (domW + 127) >> 7;
    const int nseg = nseg_row * domH;  // 128-pixel steps; loop-domain pixels < 2^31 (checked on the host)
    const int lane = (int)(threadIdx.x & 31u);
    const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
    int t_have = -1;
    for (int sg = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); sg < nseg; sg += nwarps) {
        const int yy = sg / nseg_row;
        const int xx0 = ((sg - yy * nseg_row) << 7) + lane;
        const double y = (double)(minY + yy);
        const int key0 = yy * domW + xx0;
        int tq[4] = {-1, -1, -1, -1};
        if (MODE == 2) {
            // map[key]: entries past the end read `undefined` (> -1 is false); Int16Array semantics of the stored id
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (xx0 + 32 * k < domW && (long long)key0 + 32 * k < map_len) {
                    const int raw = __ldg(map32 + key0 + 32 * k);
                    const int t = (raw < 0) ? -1 : (int)(short)(unsigned short)(raw & 0xFFFF);
                    tq[k] = (t >= 0 && t < n_tris) ? t : -1;
                }
            }
        }
        double r0 = 0.0, r1 = 0.0, r2 = 0.0;
        if (MODE == 0) {
            r0 = __dmul_rn(m[2], y);
            r1 = __dmul_rn(m[3], y);
        } else if (MODE == 1) {
            r0 = __dmul_rn(m[1], y);
            r1 = __dmul_rn(m[4], y);
            r2 = __dmul_rn(m[7], y);
        }
        const double x_first = (double)(minX + xx0);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (xx0 + 32 * k >= domW) break;
            const double x = __dadd_rn(x_first, (double)(32 * k));   // exact: small integers
            double tx, ty;
            if (MODE == 2) {
                const int t = tq[k];
                if (t < 0) continue;
                if (t != t_have) {
                    const float *f = rec[t].fwd;
#pragma unroll
                    for (int c = 0; c < 6; ++c) m[c] = (double)__ldg(f + c);
                    t_have = t;
                }
                tx = __dadd_rn(__dadd_rn(__dmul_rn(m[0], x), __dmul_rn(m[2], y)), m[4]);
                ty = __dadd_rn(__dadd_rn(__dmul_rn(m[1], x), __dmul_rn(m[3], y)), m[5]);
            } else if (MODE == 0) {
                tx = __dadd_rn(__dadd_rn(__dmul_rn(m[0], x), r0), m[4]);
                ty = __dadd_rn(__dadd_rn(__dmul_rn(m[1], x), r1), m[5]);
            } else {
                const double den = __dadd_rn(__dadd_rn(__dmul_rn(m[6], x), r2), 1.0);
                tx = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(m[0], x), r0), m[2]), den);
                ty = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(m[3], x), r1), m[5]), den);
            }
            const long long p = forward_target(tx, ty, xOff, yOff, oW, npix_out);
            if (p >= 0) atomicMax(winner + p, key0 + 32 * k);
        }
    }
}

// pass 2: every output pixel takes the source pixel of its winning key (or stays transparent) and hands the plane
// entry back as -1.  When the loop domain is the image itself (always for _geometricWarp; for _piecewiseAffineWarp when
// the source points span the image) the key IS the flat source index: no division.
#ifndef HG_FWD_GATHER_U
#define HG_FWD_GATHER_U 2
#endif
constexpr int FWD_GU = HG_FWD_GATHER_U;   // quads per thread and step of the gather pass
__global__ void __launch_bounds__(256) forward_gather_kernel(const FwdParams P)
{
    const FwdArgs &g = P.many ? P.many[blockIdx.y] : P.one;
    if (g.lattice) return;
    // the frame's constants in registers (the stores below may alias the descriptor as far as the compiler knows)
    struct { int *winner; uint32_t *out; int domW, W, minX, minY; } a = {g.winner, g.out, g.domW, g.W, g.minX, g.minY};
    const int npix = g.oW * g.oH;  // < 2^31 (checked on the host)
    const int nquad = (npix + 3) >> 2;
    const long long npx_src = (long long)g.W * g.H;
    const int stride = (int)(gridDim.x * blockDim.x);
    const bool key_is_flat = g.domW == g.W && g.minX == 0 && g.minY == 0;
    const uint32_t *__restrict__ src = g.src;
    // two quads per thread and step, a grid stride apart: both plane reads go out first, then the eight gathers that depend
    // on them, then the two stores (the pass is bound by this chain of dependent loads)
    for (int q = (int)(blockIdx.x * blockDim.x + threadIdx.x); q < nquad; q += FWD_GU * stride) {
        int key[FWD_GU][4];
        uint32_t px[FWD_GU][4];
#pragma unroll
        for (int u = 0; u < FWD_GU; ++u) {
            const int p0 = (q + u * stride) << 2;
#pragma unroll
            for (int k = 0; k < 4; ++k) key[u][k] = -1;
            if (q + u * stride >= nquad) continue;
            if (p0 + 3 < npix) {
                int4 *wp = reinterpret_cast<int4 *>(a.winner + p0);
                const int4 kv = *wp;
                key[u][0] = kv.x; key[u][1] = kv.y; key[u][2] = kv.z; key[u][3] = kv.w;
                if ((kv.x & kv.y & kv.z & kv.w) != -1) *wp = make_int4(-1, -1, -1, -1);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (p0 + k < npix) {
                        key[u][k] = a.winner[p0 + k];
                        if (key[u][k] != -1) a.winner[p0 + k] = -1;
                    }
            }
        }
#pragma unroll
        for (int u = 0; u < FWD_GU; ++u)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t v = 0u;
                if (key[u][k] >= 0) {
                    long long flat = key[u][k];
                    if (!key_is_flat) {
                        const int yy = key[u][k] / a.domW;
                        const int xx = key[u][k] - yy * a.domW;
                        // idx = y*(W<<2) + (x<<2): flat, so x >= W runs into the next row; outside the image -> 0
                        flat = (long long)(a.minY + yy) * a.W + (a.minX + xx);
                    }
                    if (flat >= 0 && flat < npx_src) v = __ldg(src + flat);
                }
                px[u][k] = v;
            }
#pragma unroll
        for (int u = 0; u < FWD_GU; ++u) {
            const int p0 = (q + u * stride) << 2;
            if (q + u * stride >= nquad) continue;
            if (p0 + 3 < npix) {
                *reinterpret_cast<uint4 *>(a.out + p0) = make_uint4(px[u][0], px[u][1], px[u][2], px[u][3]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (p0 + k < npix) a.out[p0 + k] = px[u][k];
            }
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
    // by value: through a reference every field would be re-read after each store (which may alias the descriptor)
    const FwdArgs a = P.many ? P.many[blockIdx.y] : P.one;
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
