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

// one warp per triangle, lanes stride over its rows
__global__ void __launch_bounds__(128) pw_span_bin_kernel(const FusedFrame *frames)
{
    const FusedFrame &F = frames[blockIdx.y];
    const int t = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (t >= F.n_tris) return;
    const int lane = threadIdx.x & 31;
    const TriRec &r = F.rec[t];
    const long long len = (long long)F.oW * F.oH;
    const double mw = (double)F.oW, yoff = (double)F.yOff;
    for (long long i = lane;; i += 32) {
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

constexpr int PWF_ROWS = 4;                    // rows per thread
constexpr int PWF_TY = 8;                      // thread rows per CTA
constexpr int PWF_THREADS = 16 * PWF_TY;       // 16 quads (= one 64-column bin) x 8
constexpr int PWF_TILE_ROWS = PWF_TY * PWF_ROWS;

__host__ __device__ inline int pwf_tiles_x(int oW) { return (oW + PW_BIN_W - 1) / PW_BIN_W; }
__host__ __device__ inline int pwf_tiles_y(int oH) { return (oH + PWF_TILE_ROWS - 1) / PWF_TILE_ROWS; }

// CTA = one 64-column bin x 32 rows; thread = one quad x 4 consecutive rows (16 pixels, 16 gathers in flight)
__global__ void __launch_bounds__(PWF_THREADS, 4) pw_warp_fused_kernel(const FusedFrame *frames)
{
    const FusedFrame F = frames[blockIdx.y];
    const int tiles_x = pwf_tiles_x(F.oW);
    const int tile_y = blockIdx.x / tiles_x;
    const int tile_x = blockIdx.x - tile_y * tiles_x;
    if (tile_y * PWF_TILE_ROWS >= F.oH) return;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int c_rel = tx * 4;                       // first column of the quad inside the bin
    const int xx0 = tile_x * PW_BIN_W + c_rel;      // output column
    if (xx0 >= F.oW) return;
    const int yy0 = tile_y * PWF_TILE_ROWS + ty * PWF_ROWS;
    if (yy0 >= F.oH) return;
    const uint32_t *__restrict__ src = F.src;
    const long long npx_src = (long long)F.W * F.H;
    const bool vec = ((F.oW & 3) == 0);             // dense rows are 16-byte aligned only then
    const int nvalid = min(4, F.oW - xx0);

    double xs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) xs[k] = (double)(F.xOff + xx0 + k);

    long long flat[PWF_ROWS][4];
    // the inverse matrix of the triangle the thread is currently inside: neighbouring pixels and rows mostly share
    // it, so it is (re)loaded only when the triangle id changes (3 x 16 B from L1/L2 instead of 48 B per pixel)
    int cur_t = -1;
    double m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0, m5 = 0;
#pragma unroll
    for (int j = 0; j < PWF_ROWS; ++j) {
        const int yy = yy0 + j;
        int best[4] = {-1, -1, -1, -1};
        if (yy < F.oH) {
            const size_t bin = (size_t)yy * F.bins_x + tile_x;
            const unsigned cnt = min(__ldg(F.bin_cnt + bin), (unsigned)PW_BIN_CAP);
            for (unsigned e = 0; e < cnt; ++e) {
                const unsigned ent = __ldg(F.bin_ent + bin * PW_BIN_CAP + e);
                const int lo = (int)(ent & 127u), hi = (int)((ent >> 7) & 127u), t = (int)(ent >> 14);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if ((unsigned)(c_rel + k - lo) < (unsigned)(hi - lo)) best[k] = max(best[k], t);
            }
        }
        const double y = (double)(F.yOff + yy);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            long long f = -1;
            // Int16Array semantics of the map: the stored id is t mod 2^16 as int16; negative = no triangle
            const int t = (best[k] < 0) ? -1 : (int)(short)(unsigned short)(best[k] & 0xFFFF);
            if (t >= 0 && t < F.n_tris) {
                if (t != cur_t) {
                    const double2 *m = reinterpret_cast<const double2 *>(F.inv + 6 * (size_t)t);
                    const double2 m01 = __ldg(m), m23 = __ldg(m + 1), m45 = __ldg(m + 2);
                    m0 = m01.x; m1 = m01.y; m2 = m23.x; m3 = m23.y; m4 = m45.x; m5 = m45.y;
                    cur_t = t;
                }
                const double sx = affine_coord_exact(m0, xs[k], __dmul_rn(m2, y), m4);
                const double sy = affine_coord_exact(m1, xs[k], __dmul_rn(m3, y), m5);
                const double tx2 = __dadd_rd(sx, HG_MAGIC), ty2 = __dadd_rd(sy, HG_MAGIC);
                const int ix = __double2hiint(tx2) - HG_HI_ZERO, iy = __double2hiint(ty2) - HG_HI_ZERO;
                // minSrcX <= sx < W + minSrcX and minSrcY <= sy < H + minSrcY  (H.js:1047)
                if ((unsigned)(ix - F.minSrcX) < (unsigned)F.W && (unsigned)(iy - F.minSrcY) < (unsigned)F.H) {
                    const int rx = ix + (int)((unsigned)__double2loint(tx2) >> 31);
                    const int ry = iy + (int)((unsigned)__double2loint(ty2) >> 31);
                    const long long fl = (long long)ry * F.W + rx;
                    if (fl >= 0 && fl < npx_src) f = fl;
                }
            }
            flat[j][k] = f;
        }
    }
    uint32_t px[PWF_ROWS][4];
#pragma unroll
    for (int j = 0; j < PWF_ROWS; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) px[j][k] = (flat[j][k] >= 0) ? __ldg(src + flat[j][k]) : 0u;
#pragma unroll
    for (int j = 0; j < PWF_ROWS; ++j) {
        const int yy = yy0 + j;
        if (yy < F.oH) {
            uint32_t *dst = F.out + ((long long)yy * F.oW + xx0);
            if (vec && nvalid == 4) {
                *reinterpret_cast<uint4 *>(dst) = make_uint4(px[j][0], px[j][1], px[j][2], px[j][3]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < nvalid) dst[k] = px[j][k];
            }
        }
    }
}

}  // namespace hg
