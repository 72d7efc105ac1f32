// piecewise_fused.cuh — K3/K4 fused: inverse piecewise-affine warp WITHOUT a per-pixel index map.
//
// The reference rebuilds an Int16 index map of oW*oH entries on every inverse piecewise warp (H.js:1033 ->
// 845-861) and reads it back per pixel (H.js:1044).  On a GPU that is 2-4 extra bytes per pixel in each direction
// next to the 8 algorithmic ones.  Here the map is never materialised.  fillTriangle (H.js:1111-1126) writes, for
// each triangle t and each of its rows y, ONE flat interval [S,E) of the map (TypedArray.fill semantics, computed
// exactly like the general path in piecewise.cuh); later triangles overwrite earlier ones, i.e. map[p] is the
// MAXIMUM t over the intervals that contain p.  The intervals are cut at map-row boundaries and binned:
//
//   bin (row R, 64-column block B)  ->  up to PW_BIN_CAP entries  (t, c0, c1)   "triangle t covers columns
//                                                                               [c0,c1) of this block in row R"
//
// (~1/64 of the map's size; a few entries per bin for any non-folded mesh).  The warp kernel resolves
// t = max over the matching entries of its pixel's bin, then does what H.js:1046-1052 does: inverse 2x3 of the
// triangle, window test, Math.round, flat gather, 128-bit store.
//
// Exactness: every quirk of the reference's map (no x offset -> spans spilling into the next row, negative
// relative fill indices landing at the END of the map, last-writer-wins overlaps, int16 wrap of ids) is inherited
// from the exact interval computation.  A frame that cannot be represented (a bin with more than PW_BIN_CAP
// entries, an interval crossing more than PW_MAX_PIECES rows, >= 2^17 triangles) raises a status flag and is redone
// by the general map-based path — never approximated.
#pragma once
#include "piecewise.cuh"
#include "warp_geo.cuh"

namespace hg {

constexpr int PW_BIN_W = 64;
constexpr int PW_BIN_CAP = 8;
constexpr int PW_MAX_PIECES = 6;
constexpr int PW_MAX_TRIS = 1 << 17;

struct FusedFrame {
    const uint32_t *src;
    uint32_t *out;
    const TriRec *rec;     // n_tris records of this frame (edges + row range)
    const double *inv;     // n_tris * 6 doubles: the f32-rounded inverse matrices, widened
    unsigned *bin_cnt;     // oH * bins_x counters (zeroed before the span kernel)
    unsigned *bin_ent;     // oH * bins_x * PW_BIN_CAP packed entries  (t << 14 | c1 << 7 | c0)
    int *status;           // bit 0: not representable -> redo with the general path
    int W, H, xOff, yOff, oW, oH, minSrcX, minSrcY, n_tris, bins_x;
};

// `split` warps per triangle (tall triangles of small meshes would otherwise leave the machine empty); the lanes of
// those warps stride over the triangle's rows
__global__ void __launch_bounds__(128) pw_span_bin_kernel(const FusedFrame *frames, int split)
{
    const FusedFrame &F = frames[blockIdx.y];
    const int gw = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int t = gw / split;
    if (t >= F.n_tris) return;
    const int lane = (threadIdx.x & 31) + 32 * (gw - t * split);
    const int stride = 32 * split;
    const TriRec &r = F.rec[t];
    const long long len = (long long)F.oW * F.oH;
    const double mw = (double)F.oW, yoff = (double)F.yOff;
    for (long long i = lane;; i += stride) {
        const double y = (double)r.y0 + (double)i;
        if (!(y < r.maxY)) break;  // also ends on NaN
        double xo, xd;
        predict_x_limits(r, y, xo, xd);
        const double rowbase = __dmul_rn(__dsub_rn(y, yoff), mw);
        long long k0 = js_fill_bound(__dadd_rn(rowbase, js_round(xo)), len);
        const long long k1 = js_fill_bound(__dadd_rn(rowbase, js_round(xd)), len);
        int pieces = 0;
        while (k0 < k1) {
            if (++pieces > PW_MAX_PIECES) { atomicOr(F.status, 1); break; }
            const long long row = k0 / F.oW;
            const long long row_end = (row + 1) * F.oW;
            const long long e = k1 < row_end ? k1 : row_end;
            const int c0 = (int)(k0 - row * F.oW), c1 = (int)(e - row * F.oW);
            for (int b = c0 / PW_BIN_W; b <= (c1 - 1) / PW_BIN_W; ++b) {
                const int lo = max(c0, b * PW_BIN_W) - b * PW_BIN_W;
                const int hi = min(c1, (b + 1) * PW_BIN_W) - b * PW_BIN_W;
                const size_t bin = (size_t)row * F.bins_x + b;
                const unsigned slot = atomicAdd(F.bin_cnt + bin, 1u);
                if (slot < PW_BIN_CAP) F.bin_ent[bin * PW_BIN_CAP + slot] = ((unsigned)t << 14) | ((unsigned)hi << 7) | (unsigned)lo;
                else atomicOr(F.status, 1);
            }
            k0 = e;
        }
    }
}

#ifndef HG_PWF_MINB
#define HG_PWF_MINB 5
#endif
constexpr int PWF_TX = 8;                       // threads across one 64-column bin: each owns quads tx and tx+8
constexpr int PWF_TY = 16;                      // thread rows per CTA
constexpr int PWF_THREADS = PWF_TX * PWF_TY;    // 128
constexpr int PWF_GROUP_ROWS = PWF_TY;          // rows one CTA covers per iteration (one row per thread)

__host__ __device__ inline int pwf_tiles_x(int oW) { return (oW + PW_BIN_W - 1) / PW_BIN_W; }
__host__ __device__ inline int pwf_tiles_y(int oH, int niter) { return (oH + PWF_GROUP_ROWS * niter - 1) / (PWF_GROUP_ROWS * niter); }

__device__ __forceinline__ void pwf_load_matrix(const double *inv, int t, double (&m)[6])
{
    const double2 *p = reinterpret_cast<const double2 *>(inv + 6 * (size_t)t);
    const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    m[0] = a.x; m[1] = a.y; m[2] = b.x; m[3] = b.y; m[4] = c.x; m[5] = c.y;
}

// coordinates -> flat source index or HG_OUTSIDE: window test [minSrc, W+minSrc) x [minSrc, H+minSrc) on the unrounded
// coordinate (H.js:1047), Math.round, flat index, reads outside the image give nothing (H.js:1048-1052).
// ZERO_OFF: minSrcX == minSrcY == 0 (source points start at the image origin — the usual case): everything fits
// unsigned 32-bit arithmetic exactly like the affine / projective kernels.
template <bool ZERO_OFF>
__device__ __forceinline__ unsigned pwf_decode(double sx, double sy, const FusedFrame &F, unsigned npx_src)
{
    const double tx2 = __dadd_rd(sx, HG_MAGIC), ty2 = __dadd_rd(sy, HG_MAGIC);
    if (ZERO_OFF) return decode_flat(tx2, ty2, (unsigned)F.W, (unsigned)F.H, npx_src);
    const int ix = __double2hiint(tx2) - HG_HI_ZERO, iy = __double2hiint(ty2) - HG_HI_ZERO;
    const int rx = ix + (int)((unsigned)__double2loint(tx2) >> 31);
    const int ry = iy + (int)((unsigned)__double2loint(ty2) >> 31);
    const long long fl = (long long)ry * F.W + rx;
    const bool ok = ((unsigned)(ix - F.minSrcX) < (unsigned)F.W) & ((unsigned)(iy - F.minSrcY) < (unsigned)F.H) &
                    (fl >= 0) & (fl < (long long)npx_src);
    return ok ? (unsigned)fl : HG_OUTSIDE;
}

// Int16Array semantics of the map (H.js:848): the stored id is t mod 2^16 as int16, negative = no triangle; ids that
// would index past the matrix list never occur for maps built from the same mesh
__device__ __forceinline__ int pwf_map_id(int raw, int n_tris)
{
    const int t = (raw < 0) ? -1 : (int)(short)(unsigned short)(raw & 0xFFFF);
    return (t >= 0 && t < n_tris) ? t : -1;
}

// CTA = one 64-column bin column x (16 * niter) rows.  Thread (tx,ty) owns ONE row per row group and TWO quads of
// it: columns 4tx..4tx+3 and 32+4tx..32+4tx+3 — both inside the same bin, so the bin's entries are decoded once per
// 8 pixels, and a warp (8 lanes x 4 rows) stores 2 x 128 contiguous bytes per row.
// FIVE-STAGE SOFTWARE PIPELINE inside each warp — every stage consumes what an earlier iteration produced, so the
// three dependent memory round trips of the path (bins -> triangle matrix -> source pixel) overlap with arithmetic:
//   iteration i:  S4 store row group i-4 | S3 issue gathers of group i-3 | S2 coordinates of group i-2 (H.js:1046-1048)
//                 S1 resolve triangle ids of group i-1 from its bin, issue the matrix loads | S0 issue bin loads of i
// Each quad keeps the inverse matrix of its first pixel's triangle in registers; a quad that a span boundary cuts
// through resolves its pixels one by one (and fetches a second matrix on the spot).
template <bool ZERO_OFF>
__device__ __forceinline__ void pwf_body(const FusedFrame &F, int niter, int tile_x, int row0)
{
    const int tx = threadIdx.x & (PWF_TX - 1), ty = threadIdx.x / PWF_TX;
    const int base0 = row0 + ty;
    if (base0 >= F.oH) return;
    const int c_rel[2] = {tx * 4, 32 + tx * 4};                          // first column of each quad inside the bin
    const int xx0[2] = {tile_x * PW_BIN_W + c_rel[0], tile_x * PW_BIN_W + c_rel[1]};  // output columns
    if (xx0[0] >= F.oW) return;
    const int ngroups = min(niter, (F.oH - base0 + PWF_GROUP_ROWS - 1) / PWF_GROUP_ROWS);
    const uint32_t *__restrict__ src = F.src;
    const unsigned npx_src = (unsigned)F.W * (unsigned)F.H;
    const bool aligned = (F.oW & 3) == 0;  // dense rows are 16-byte aligned only when oW % 4 == 0
    int nvalid[2];
    nvalid[0] = min(4, F.oW - xx0[0]);
    nvalid[1] = max(0, min(4, F.oW - xx0[1]));
    const int n_tris = F.n_tris;

    double xs[2][4];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k) xs[q][k] = (double)(F.xOff + xx0[q] + k);

    const unsigned *p_cnt = F.bin_cnt + ((size_t)base0 * F.bins_x + tile_x);
    const uint4 *p_ent = reinterpret_cast<const uint4 *>(F.bin_ent) + 2 * ((size_t)base0 * F.bins_x + tile_x);
    const size_t cnt_step = (size_t)PWF_GROUP_ROWS * F.bins_x;
    uint32_t *p_out = F.out + ((long long)base0 * F.oW + xx0[0]);
    const long long out_step = (long long)PWF_GROUP_ROWS * F.oW;

    unsigned bcnt = 0;
    uint4 be0 = make_uint4(0, 0, 0, 0), be1 = be0;
    int tri[2][4];
    unsigned uni = 0;      // bit q: the four pixels of quad q share one triangle id
    double mq[2][6];       // inverse matrix of the first pixel's triangle of each quad
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int k = 0; k < 6; ++k) mq[q][k] = 0.0;
    unsigned idx[2][4];
    uint32_t px[2][4];

#pragma unroll 1
    for (int it = 0; it < ngroups + 4; ++it) {
        // ---- S4: store row group it-4
        if (it >= 4) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                uint32_t *dst = p_out + 32 * q;
                if (aligned && nvalid[q] == 4) {
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(px[q][0], px[q][1], px[q][2], px[q][3]);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < nvalid[q]) dst[k] = px[q][k];
                }
            }
            p_out += out_step;
        }
        // ---- S3: gathers of group it-3
        if (it >= 3 && it - 3 < ngroups) {
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int k = 0; k < 4; ++k) px[q][k] = ldg_or_zero(src, idx[q][k]);
        }
        // ---- S2: coordinates of group it-2
        if (it >= 2 && it - 2 < ngroups) {
            const double y = (double)(F.yOff + base0 + (it - 2) * PWF_GROUP_ROWS);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int t0 = tri[q][0];
                if ((uni >> q) & 1u) {
                    // common case: the four pixels of the quad lie in one triangle, whose matrix is in mq[q]
                    if (t0 >= 0) {
                        const double r0 = __dmul_rn(mq[q][2], y), r1 = __dmul_rn(mq[q][3], y);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            idx[q][k] = pwf_decode<ZERO_OFF>(affine_coord_exact(mq[q][0], xs[q][k], r0, mq[q][4]),
                                                             affine_coord_exact(mq[q][1], xs[q][k], r1, mq[q][5]), F, npx_src);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) idx[q][k] = HG_OUTSIDE;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int t = tri[q][k];
                        unsigned f = HG_OUTSIDE;
                        if (t >= 0) {
                            double m[6];
                            if (t != t0) {
                                pwf_load_matrix(F.inv, t, m);  // a span boundary cuts through the quad
                            } else {
#pragma unroll
                                for (int c = 0; c < 6; ++c) m[c] = mq[q][c];
                            }
                            f = pwf_decode<ZERO_OFF>(affine_coord_exact(m[0], xs[q][k], __dmul_rn(m[2], y), m[4]),
                                                     affine_coord_exact(m[1], xs[q][k], __dmul_rn(m[3], y), m[5]), F, npx_src);
                        }
                        idx[q][k] = f;
                    }
                }
            }
        }
        // ---- S1: triangle ids of group it-1 from its bin; fetch the matrices S2 will need next iteration
        if (it >= 1 && it - 1 < ngroups) {
            int tfull[2] = {-1, -1};                                   // highest id among entries covering a WHOLE quad
            int best[2][4] = {{-1, -1, -1, -1}, {-1, -1, -1, -1}};     // per pixel, for entries cutting through a quad
            unsigned mixed = 0;
            const unsigned cnt = min(bcnt, (unsigned)PW_BIN_CAP);
            const unsigned ent[8] = {be0.x, be0.y, be0.z, be0.w, be1.x, be1.y, be1.z, be1.w};
#pragma unroll
            for (int e = 0; e < PW_BIN_CAP; ++e) {
                if ((unsigned)e >= cnt) break;
                const int lo = (int)(ent[e] & 127u), hi = (int)((ent[e] >> 7) & 127u), t = (int)(ent[e] >> 14);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (lo <= c_rel[q] && c_rel[q] + 4 <= hi) {
                        tfull[q] = max(tfull[q], t);
                    } else if (lo < c_rel[q] + 4 && c_rel[q] < hi) {
                        mixed |= 1u << q;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if ((unsigned)(c_rel[q] + k - lo) < (unsigned)(hi - lo)) best[q][k] = max(best[q][k], t);
                    }
                }
            }
            uni = 0;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (!((mixed >> q) & 1u)) {
                    const int t = pwf_map_id(tfull[q], n_tris);
#pragma unroll
                    for (int k = 0; k < 4; ++k) tri[q][k] = t;
                    uni |= 1u << q;
                } else {
                    bool u = true;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        tri[q][k] = pwf_map_id(max(tfull[q], best[q][k]), n_tris);
                        u = u && (tri[q][k] == tri[q][0]);
                    }
                    uni |= u ? (1u << q) : 0u;
                }
                if (tri[q][0] >= 0) pwf_load_matrix(F.inv, tri[q][0], mq[q]);
            }
        }
        // ---- S0: bin loads of group it
        if (it < ngroups) {
            bcnt = __ldg(p_cnt);
            be0 = __ldg(p_ent);
            be1 = __ldg(p_ent + 1);
            p_cnt += cnt_step;
            p_ent += 2 * cnt_step;
        }
    }
}

__global__ void __launch_bounds__(PWF_THREADS, HG_PWF_MINB) pw_warp_fused_kernel(const FusedFrame *frames, int niter)
{
    const FusedFrame F = frames[blockIdx.y];
    const int tiles_x = pwf_tiles_x(F.oW);
    const int tile_y = blockIdx.x / tiles_x;
    const int tile_x = blockIdx.x - tile_y * tiles_x;
    const int row0 = tile_y * PWF_GROUP_ROWS * niter;
    if (row0 >= F.oH) return;
    if (F.minSrcX == 0 && F.minSrcY == 0) pwf_body<true>(F, niter, tile_x, row0);
    else pwf_body<false>(F, niter, tile_x, row0);
}

}  // namespace hg
