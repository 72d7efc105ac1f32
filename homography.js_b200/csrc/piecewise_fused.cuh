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
// (~1/64 of the map's size; a few entries per bin for any non-folded mesh).  A second pass (pw_bin_runs_kernel, one
// thread per bin) resolves the overlaps — highest id wins — and run-length encodes each bin as a 64-bit start mask +
// up to 8 int16 ids.  The warp kernel finds t with two popcounts, then does what H.js:1046-1052 does: inverse 2x3 of
// the triangle, window test, Math.round, flat gather, 128-bit store.
//
// Exactness: every quirk of the reference's map (no x offset -> spans spilling into the next row, negative
// relative fill indices landing at the END of the map, last-writer-wins overlaps, int16 wrap of ids) is inherited
// from the exact interval computation.  A frame that cannot be represented (a bin with more than PW_BIN_CAP
// entries or runs, an interval crossing more than PW_MAX_PIECES rows, >= 2^17 triangles) raises a status flag and is redone
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
    uint4 *bin_run;        // oH * bins_x run records (pw_bin_runs_kernel): what the pixel kernel reads
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
            // eight bins at a time: the slot reservations (independent atomics) go out back to back, the entry
            // stores that depend on them follow
            const int b_last = (c1 - 1) / PW_BIN_W;
            for (int b0 = c0 / PW_BIN_W; b0 <= b_last; b0 += 8) {
                unsigned slot[8];
                const size_t bin0 = (size_t)row * F.bins_x + b0;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (b0 + i <= b_last) slot[i] = atomicAdd(F.bin_cnt + bin0 + i, 1u);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (b0 + i <= b_last) {
                        const int b = b0 + i;
                        const int lo = max(c0, b * PW_BIN_W) - b * PW_BIN_W;
                        const int hi = min(c1, (b + 1) * PW_BIN_W) - b * PW_BIN_W;
                        if (slot[i] < PW_BIN_CAP) F.bin_ent[(bin0 + i) * PW_BIN_CAP + slot[i]] = ((unsigned)t << 14) | ((unsigned)hi << 7) | (unsigned)lo;
                        else atomicOr(F.status, 1);
                    }
                }
            }
            k0 = e;
        }
    }
}


// Int16Array semantics of the map (H.js:848): the stored id is t mod 2^16 as int16, negative = no triangle; ids that
// would index past the matrix list never occur for maps built from the same mesh
__device__ __forceinline__ int pwf_map_id(int raw, int n_tris)
{
    const int t = (raw < 0) ? -1 : (int)(short)(unsigned short)(raw & 0xFFFF);
    return (t >= 0 && t < n_tris) ? t : -1;
}

// Second binning pass, one thread per bin: resolve the overlaps ONCE (highest triangle id wins, then the Int16 wrap)
// and run-length encode the 64 columns of the bin:
//     record = { start mask (bit c: a run begins at column c), up to PW_RUN_CAP ids as int16, in column order }
// so the pixel kernel finds a pixel's triangle with two popcounts instead of scanning the entries: run index =
// popc(mask & bits[0..c]) - 1.  Adjacent runs with the same final id are merged; more than PW_RUN_CAP runs in one bin
// flags the frame for the general path.
constexpr int PW_RUN_CAP = 8;

__global__ void __launch_bounds__(128) pw_bin_runs_kernel(const FusedFrame *frames)
{
    const FusedFrame &F = frames[blockIdx.y];
    const size_t nbins = (size_t)F.bins_x * F.oH;
    const size_t bin = (size_t)blockIdx.x * 128 + threadIdx.x;
    if (bin >= nbins) return;
    const unsigned cnt = min(F.bin_cnt[bin], (unsigned)PW_BIN_CAP);
    unsigned ent[PW_BIN_CAP];
    {
        const uint4 *pe = reinterpret_cast<const uint4 *>(F.bin_ent) + 2 * bin;
        const uint4 a = cnt > 0 ? pe[0] : make_uint4(0, 0, 0, 0), b = cnt > 4 ? pe[1] : make_uint4(0, 0, 0, 0);
        ent[0] = a.x; ent[1] = a.y; ent[2] = a.z; ent[3] = a.w; ent[4] = b.x; ent[5] = b.y; ent[6] = b.z; ent[7] = b.w;
    }
    // candidate run starts: column 0 and every interval end point inside the bin.  Typical bins hold 1-3 entries: the
    // loops stop at cnt instead of running predicated over all PW_BIN_CAP slots
    unsigned long long cand = 1ull;
#pragma unroll
    for (int e = 0; e < PW_BIN_CAP; ++e) {
        if ((unsigned)e >= cnt) break;
        const unsigned lo = ent[e] & 127u, hi = (ent[e] >> 7) & 127u;
        cand |= 1ull << lo;          // lo <= 63
        if (hi < 64u) cand |= 1ull << hi;
    }
    unsigned long long mask = 0ull;
    unsigned ids[PW_RUN_CAP / 2] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};  // int16 pairs, -1 = no triangle
    int runs = 0, prev = -2;
    while (cand) {
        const int c = __ffsll((long long)cand) - 1;
        cand &= cand - 1;
        int raw = -1;
#pragma unroll
        for (int e = 0; e < PW_BIN_CAP; ++e) {
            if ((unsigned)e >= cnt) break;
            // lo <= c < hi  <=>  (unsigned)(c - lo) < (unsigned)(hi - lo)
            const int lo = (int)(ent[e] & 127u), hi = (int)((ent[e] >> 7) & 127u);
            if ((unsigned)(c - lo) < (unsigned)(hi - lo)) raw = max(raw, (int)(ent[e] >> 14));
        }
        const int id = pwf_map_id(raw, F.n_tris);
        if (id != prev) {
            if (runs == PW_RUN_CAP) {
                atomicOr(F.status, 1);
                break;
            }
            const unsigned h = (unsigned)id & 0xFFFFu;
#pragma unroll
            for (int w = 0; w < PW_RUN_CAP / 2; ++w)
                if (w == (runs >> 1)) ids[w] = (runs & 1) ? ((ids[w] & 0x0000FFFFu) | (h << 16)) : ((ids[w] & 0xFFFF0000u) | h);
            mask |= 1ull << c;
            prev = id;
            ++runs;
        }
    }
    F.bin_run[2 * bin] = make_uint4((unsigned)mask, (unsigned)(mask >> 32), 0u, 0u);
    F.bin_run[2 * bin + 1] = make_uint4(ids[0], ids[1], ids[2], ids[3]);
}

// id of run r (0..7) from the packed int16 ids
__device__ __forceinline__ int pwf_run_id(const uint4 &ids, unsigned r)
{
    const unsigned lo = (r & 2u) ? ids.y : ids.x, hi = (r & 2u) ? ids.w : ids.z;
    const unsigned w = (r & 4u) ? hi : lo;
    return (int)(short)(unsigned short)((r & 1u) ? (w >> 16) : w);
}

#ifndef HG_PWF_MINB
#define HG_PWF_MINB 5
#endif
constexpr int PWF_TX = 8;                       // threads across one 64-column bin: each owns quads tx and tx+8
constexpr int PWF_TY = 16;                      // thread rows per CTA
constexpr int PWF_THREADS = PWF_TX * PWF_TY;    // 128
constexpr int PWF_GROUP_ROWS = PWF_TY;          // rows one CTA covers per iteration (one row per thread)

// 64-column bins of a map row (what the span / run kernels fill) ...
__host__ __device__ inline int pwf_bins_x(int oW) { return (oW + PW_BIN_W - 1) / PW_BIN_W; }
// ... and CTA tiles of an output row.  The pixel kernel walks rows in FLAT-aligned quads (like warp_geo.cuh): row yy is
// shifted left by a(yy) = (yy * oW) & 3 columns, so when oW % 4 != 0 its last pixels may sit up to 3 columns further right
__host__ __device__ inline int pwf_tiles_x(int oW) { return ((oW & 3) ? (oW + 3 + PW_BIN_W - 1) : (oW + PW_BIN_W - 1)) / PW_BIN_W; }
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
    if (ZERO_OFF) {
        // doubled coordinates (see geo_fast_body): n = hi(2v + 1 + magic) - HG_HI_ZERO = floor(2v + 1), exact for every v;
        // 0 <= v < W  <=>  (unsigned)(n - 1) < 2W,  Math.round(v) = n >> 1
        const unsigned hx = (unsigned)__double2hiint(__fma_rd(sx, 2.0, HG_MAGIC + 1.0));
        const unsigned hy = (unsigned)__double2hiint(__fma_rd(sy, 2.0, HG_MAGIC + 1.0));
        const unsigned flat = (hy >> 1) * (unsigned)F.W + (hx >> 1) - (unsigned)(HG_HI_ZERO >> 1) * ((unsigned)F.W + 1u);
        const bool in = ((hx - (unsigned)(HG_HI_ZERO + 1)) < 2u * (unsigned)F.W) & ((hy - (unsigned)(HG_HI_ZERO + 1)) < 2u * (unsigned)F.H) &
                        (flat < npx_src);
        return in ? flat : HG_OUTSIDE;
    }
    const double tx2 = __dadd_rd(sx, HG_MAGIC), ty2 = __dadd_rd(sy, HG_MAGIC);
    const int ix = __double2hiint(tx2) - HG_HI_ZERO, iy = __double2hiint(ty2) - HG_HI_ZERO;
    const int rx = ix + (int)((unsigned)__double2loint(tx2) >> 31);
    const int ry = iy + (int)((unsigned)__double2loint(ty2) >> 31);
    const long long fl = (long long)ry * F.W + rx;
    const bool ok = ((unsigned)(ix - F.minSrcX) < (unsigned)F.W) & ((unsigned)(iy - F.minSrcY) < (unsigned)F.H) &
                    (fl >= 0) & (fl < (long long)npx_src);
    return ok ? (unsigned)fl : HG_OUTSIDE;
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

    double xs[2][4];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k) xs[q][k] = (double)(F.xOff + xx0[q] + k);

    const uint4 *p_run = F.bin_run + 2 * ((size_t)base0 * F.bins_x + tile_x);
    const size_t cnt_step = (size_t)PWF_GROUP_ROWS * F.bins_x;
    uint32_t *p_out = F.out + ((long long)base0 * F.oW + xx0[0]);
    const long long out_step = (long long)PWF_GROUP_ROWS * F.oW;

    uint4 be0 = make_uint4(0, 0, 0, 0), be1 = be0;  // run record of the row's bin: start mask, packed ids
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
            // quad 0 lives in the low word of the start mask (columns 0..31), quad 1 in the high word
            const unsigned below1 = __popc(be0.x);
            uni = 0;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const unsigned word = q ? be0.y : be0.x, cb = 4u * (unsigned)tx;
                const unsigned r = (q ? below1 : 0u) + __popc(word & ((2u << cb) - 1u)) - 1u;  // run of the quad's first column
                const unsigned inner = (word >> (cb + 1u)) & 7u;                               // runs starting inside the quad
                const int t0 = pwf_run_id(be1, r);
                tri[q][0] = t0;
                if (inner == 0u) {
#pragma unroll
                    for (int k = 1; k < 4; ++k) tri[q][k] = t0;
                    uni |= 1u << q;
                } else {
                    bool u = true;
#pragma unroll
                    for (int k = 1; k < 4; ++k) {
                        tri[q][k] = pwf_run_id(be1, r + __popc(inner & ((1u << k) - 1u)));
                        u = u && (tri[q][k] == t0);
                    }
                    uni |= u ? (1u << q) : 0u;
                }
                if (t0 >= 0) pwf_load_matrix(F.inv, t0, mq[q]);
            }
        }
        // ---- S0: run record of group it
        if (it < ngroups) {
            be0 = __ldg(p_run);
            be1 = __ldg(p_run + 1);
            p_run += 2 * cnt_step;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Second-generation pixel loop (the default).  Same decomposition — CTA = one 64-column tile x (16 * niter) rows, thread
// = one row x two quads per row group — with four changes measured against the first generation (profiles/r02_*):
//   * FLAT-ALIGNED QUADS.  Row yy is walked in quads that start at column 4i - a(yy), a(yy) = (yy * oW) & 3, so every
//     interior quad is one aligned 128-bit store for ANY output width (per-frame video windows have arbitrary widths);
//     16 | PWF_GROUP_ROWS, so a(yy) is the same for every row a thread owns.
//   * ONE TRIANGLE PER QUAD in the loop.  A quad is computed with the triangle of its first pixel; quads a span boundary
//     cuts through (a run starts inside them) or that reach back into the previous bin are ALSO noted in a per-warp
//     queue (slot = popcount of a ballot: no atomics) and redone after the loop, one PIXEL per lane with the pixel's own
//     triangle — dense lanes instead of a per-pixel path that every warp of a fine mesh had to walk through.
//   * the doubled-coordinate decode of warp_geo.cuh with its warp-uniform "end pixels inside => quad inside" shortcut
//     (both source coordinates are monotone along the quad: one triangle, one affine map), loads predicated directly.
//   * run ids through one PRMT; the gathers of a quad carry no per-pixel predicate when its whole warp reads inside;
//     pixel registers double-buffered (stores two iterations behind their gathers) and run records L2-prefetched: with
//     ~30 % fewer instructions than the first generation the loop was latency-bound (ncu: issue-active 47 %, the store
//     stage waiting on its gathers) until two row groups of gathers were kept in flight per thread.
constexpr int PWF_QCAP = 1024;
#ifndef PWF_PREFETCH
#define PWF_PREFETCH 3   // run records are L2-prefetched this many row groups ahead
#endif  // per warp: 16 row groups x 32 lanes x 2 quads, the most a CTA can ever queue

// id of run r (0..7) from the eight packed int16 ids: one byte permute, sign-extended
__device__ __forceinline__ int pwf_run_id2(const uint4 &ids, unsigned r)
{
    const unsigned lo = (r & 4u) ? ids.z : ids.x, hi = (r & 4u) ? ids.w : ids.y;
    // bytes {2k, 2k+1} of {lo, hi}, the upper half filled with the sign of byte 2k+1 (selector bit 3 = replicate the sign;
    // __byte_perm() masks that bit away, hence the PTX form)
    unsigned d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"((r & 3u) * 0x2222u + 0x9910u));
    return (int)d;
}

// per-thread constants and pipeline registers of pwf_body2
template <bool ZERO_OFF>
struct PwfCtx {
    const uint32_t *src;
    const double *inv;
    unsigned W, npx_src, W2, H2, Wi, Hi, kflat;
    int oW, oH, yOff, base0;
    double xs[2][4];
    unsigned vmask[2];
    unsigned long long below[2], inner[2];
    bool prev_bin, has_bin;
    int lane;
    unsigned lt_mask;
    // pipeline state
    uint4 be0, be1;      // run record (S0 -> S1)
    int t0[2];           // triangle of each quad's first pixel (S1 -> S2)
    double mq[2][6];     // its inverse matrix (S1 -> S2)
    int qn;              // warp-uniform: entries in this warp's queue
};

// S2: coordinates (H.js:1046), window test and Math.round (H.js:1047-1048), flat gather (H.js:1049-1052) of one row group
template <bool ZERO_OFF>
__device__ __forceinline__ void pwf_issue(PwfCtx<ZERO_OFF> &C, const FusedFrame &F, int g, uint32_t (&px)[2][4])
{
    const int row = C.base0 + g * PWF_GROUP_ROWS;
    const double y = (double)(C.yOff + row);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const bool live = (C.t0[q] >= 0) && (row < C.oH) && (C.vmask[q] != 0u);
        const double r0 = __dmul_rn(C.mq[q][2], y), r1 = __dmul_rn(C.mq[q][3], y);
        if (ZERO_OFF) {
            unsigned hx[4], hy[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                hx[k] = (unsigned)__double2hiint(__fma_rd(affine_coord_exact(C.mq[q][0], C.xs[q][k], r0, C.mq[q][4]), 2.0, HG_MAGIC + 1.0));
                hy[k] = (unsigned)__double2hiint(__fma_rd(affine_coord_exact(C.mq[q][1], C.xs[q][k], r1, C.mq[q][5]), 2.0, HG_MAGIC + 1.0));
            }
            const unsigned cz = (unsigned)(HG_HI_ZERO + 2);
            const bool ends_inside = live && ((hx[0] - cz) < C.Wi) & ((hy[0] - cz) < C.Hi) & ((hx[3] - cz) < C.Wi) & ((hy[3] - cz) < C.Hi);
            if (__all_sync(0xffffffffu, ends_inside)) {
#pragma unroll
                for (int k = 0; k < 4; ++k) px[q][k] = __ldg(C.src + ((hy[k] >> 1) * C.W + (hx[k] >> 1) - C.kflat));
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned flat = (hy[k] >> 1) * C.W + (hx[k] >> 1) - C.kflat;
                    const bool in = live & ((hx[k] - (unsigned)(HG_HI_ZERO + 1)) < C.W2) & ((hy[k] - (unsigned)(HG_HI_ZERO + 1)) < C.H2) &
                                    (flat < C.npx_src);
                    uint32_t v = 0u;
                    if (in) v = __ldg(C.src + flat);
                    px[q][k] = v;
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                unsigned f = HG_OUTSIDE;
                if (live)
                    f = pwf_decode<false>(affine_coord_exact(C.mq[q][0], C.xs[q][k], r0, C.mq[q][4]),
                                          affine_coord_exact(C.mq[q][1], C.xs[q][k], r1, C.mq[q][5]), F, C.npx_src);
                px[q][k] = ldg_or_zero(C.src, f);
            }
        }
    }
}

// S1: triangle of each quad's first pixel from the run record of row group g, its matrix; note cut quads in the queue
template <bool ZERO_OFF>
__device__ __forceinline__ void pwf_resolve(PwfCtx<ZERO_OFF> &C, int g, unsigned short *wq)
{
    const bool row_ok = C.base0 + g * PWF_GROUP_ROWS < C.oH;
    const unsigned long long m64 = ((unsigned long long)C.be0.y << 32) | (unsigned long long)C.be0.x;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const unsigned r = (unsigned)__popcll(m64 & C.below[q]) - 1u;
        const int t = pwf_run_id2(C.be1, r);
        C.t0[q] = t;
        if (t >= 0) pwf_load_matrix(C.inv, t, C.mq[q]);
        const bool cut = row_ok && (C.vmask[q] != 0u) && (((m64 & C.inner[q]) != 0ull) || (q == 0 && C.prev_bin));
        const unsigned bal = __ballot_sync(0xffffffffu, cut);
        if (cut) wq[C.qn + __popc(bal & C.lt_mask)] = (unsigned short)((g << 6) | (C.lane << 1) | q);
        C.qn += __popc(bal);
    }
}

// S3: stores of one row group
template <bool ZERO_OFF>
__device__ __forceinline__ void pwf_retire(const PwfCtx<ZERO_OFF> &C, int g, uint32_t *p_out, const uint32_t (&px)[2][4])
{
    if (C.base0 + g * PWF_GROUP_ROWS >= C.oH) return;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        uint32_t *dst = p_out + 32 * q;
        if (C.vmask[q] == 0xFu) {
            *reinterpret_cast<uint4 *>(dst) = make_uint4(px[q][0], px[q][1], px[q][2], px[q][3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (C.vmask[q] & (1u << k)) dst[k] = px[q][k];
        }
    }
}

template <bool ZERO_OFF>
__device__ __forceinline__ void pwf_body2(const FusedFrame &F, int niter, int tile_x, int row0, unsigned short *wq)
{
    const int warp_id = threadIdx.x >> 5;
    const int tx = threadIdx.x & (PWF_TX - 1), ty = threadIdx.x / PWF_TX;
    PwfCtx<ZERO_OFF> C;
    C.lane = threadIdx.x & 31;
    C.lt_mask = (1u << C.lane) - 1u;
    C.src = F.src;
    C.inv = F.inv;
    C.oW = F.oW;
    C.oH = F.oH;
    C.yOff = F.yOff;
    C.base0 = row0 + ty;
    const int oW = F.oW, oH = F.oH;
    const int ngroups = min(niter, (oH - row0 + PWF_GROUP_ROWS - 1) / PWF_GROUP_ROWS);  // CTA-uniform
    C.has_bin = tile_x < F.bins_x;
    const int a = (oW & 3) ? (int)(((unsigned)C.base0 * (unsigned)oW) & 3u) : 0;
    C.W = (unsigned)F.W;
    const unsigned H = (unsigned)F.H;
    C.npx_src = C.W * H;
    C.W2 = 2u * C.W;
    C.H2 = 2u * H;
    C.Wi = C.W2 >= 3u ? C.W2 - 3u : 0u;
    C.Hi = C.H2 >= 3u ? C.H2 - 3u : 0u;
    C.kflat = (unsigned)(HG_HI_ZERO >> 1) * (C.W + 1u);
    int xx00 = 0;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int c = 4 * tx + 32 * q - a;        // first column of the quad inside its bin (negative: previous bin)
        const int xx0 = tile_x * PW_BIN_W + c;
        if (q == 0) xx00 = xx0;
        C.vmask[q] = 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (xx0 + k >= 0 && xx0 + k < oW) C.vmask[q] |= 1u << k;
            C.xs[q][k] = (double)(F.xOff + xx0 + k);
        }
        const int c0 = c < 0 ? 0 : c, c3 = c + 3;                       // own-bin columns of the quad: c0 .. c3 (<= 63)
        C.below[q] = (2ull << c0) - 1ull;                               // bits 0 .. c0
        C.inner[q] = ((c3 >= 63 ? 0ull : (1ull << (c3 + 1))) - 1ull) & ~C.below[q];  // bits c0+1 .. c3
    }
    // pixels of quad 0 left of the tile belong to the previous bin (only when they are pixels of the image at all)
    C.prev_bin = (a > 0) && (tx == 0) && (tile_x > 0);
    C.be0 = make_uint4(1u, 0u, 0u, 0u);
    C.be1 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    C.t0[0] = C.t0[1] = -1;
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int k = 0; k < 6; ++k) C.mq[q][k] = 0.0;
    C.qn = 0;

    const uint4 *p_run = F.bin_run + 2 * ((size_t)C.base0 * F.bins_x + tile_x);
    const size_t run_step = 2 * (size_t)PWF_GROUP_ROWS * F.bins_x;
    uint32_t *p_out = F.out + ((long long)C.base0 * oW + xx00);
    const long long out_step = (long long)PWF_GROUP_ROWS * oW;

    // S0: the run record of row group g (L2-prefetched PWF_PREFETCH groups earlier: the records of a frame were written
    // by the run kernel just before and mostly sit in DRAM by now)
    auto load_record = [&](int g) {
        if (g >= ngroups) return;
        if (C.has_bin && C.base0 + g * PWF_GROUP_ROWS < oH) {
            C.be0 = __ldg(p_run);
            C.be1 = __ldg(p_run + 1);
            if (g + PWF_PREFETCH < ngroups && C.base0 + (g + PWF_PREFETCH) * PWF_GROUP_ROWS < oH && tx == 0)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p_run + PWF_PREFETCH * run_step));
        } else {
            C.be0 = make_uint4(1u, 0u, 0u, 0u);
            C.be1 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        }
        p_run += run_step;
    };
    if (C.has_bin && tx == 0) {
#pragma unroll
        for (int g = 1; g < PWF_PREFETCH; ++g)
            if (g < ngroups && C.base0 + g * PWF_GROUP_ROWS < oH) asm volatile("prefetch.global.L2 [%0];" ::"l"(p_run + g * run_step));
    }

    // FOUR STAGES, the pixel registers double-buffered (the loop is unrolled by two so that pa / pb alternate without
    // moves): the gathers of a row group are stored TWO iterations after they were issued, so two groups of gathers
    // (16 loads per thread) are in flight while the arithmetic of the following groups runs.
    //   iteration g:  S3 store group g-4 | S2 coordinates + gathers of g-2 | S1 ids + matrix loads of g-1 | S0 record of g
    uint32_t pa[2][4], pb[2][4];
#pragma unroll 1
    for (int it = 0; it < ngroups + 4; it += 2) {
        // even step: group it-2 gathers into pa, which group it-4 has just left
        if (it >= 4) {
            pwf_retire(C, it - 4, p_out, pa);
            p_out += out_step;
        }
        if (it >= 2 && it - 2 < ngroups) pwf_issue(C, F, it - 2, pa);
        if (it >= 1 && it - 1 < ngroups) pwf_resolve(C, it - 1, wq);
        load_record(it);
        // odd step: the same with pb
        const int i1 = it + 1;
        if (i1 >= 4 && i1 - 4 < ngroups) {
            pwf_retire(C, i1 - 4, p_out, pb);
            p_out += out_step;
        }
        if (i1 >= 2 && i1 - 2 < ngroups) pwf_issue(C, F, i1 - 2, pb);
        if (i1 - 1 < ngroups) pwf_resolve(C, i1 - 1, wq);
        load_record(i1);
    }

    // ---- the queued quads again, one QUAD per lane, every pixel with the triangle of its own run (H.js:1044-1052
    // verbatim): the run record(s) once, then up to four matrices, four gathers and one 128-bit store
    __syncwarp();  // also orders the provisional stores above before the final ones below
    const int qn = C.qn;
    for (int e0 = 0; e0 < qn; e0 += 32) {
        const int e = e0 + C.lane;
        if (e >= qn) continue;
        const unsigned ent = wq[e];
        const int qq = (int)(ent & 1u), sl = (int)((ent >> 1) & 31u), g = (int)(ent >> 6);
        const int row = row0 + warp_id * 4 + (sl >> 3) + g * PWF_GROUP_ROWS;
        const int ae = (oW & 3) ? (int)(((unsigned)row * (unsigned)oW) & 3u) : 0;
        const int X0 = tile_x * PW_BIN_W + 4 * (sl & 7) + 32 * qq - ae;
        const double y = (double)(F.yOff + row);
        // the quad lies in one bin, except a first quad that reaches back into the previous one
        const int binA = X0 < 0 ? 0 : (X0 >> 6), binB = (X0 + 3) >> 6;
        const uint4 *pr = F.bin_run + 2 * ((size_t)row * F.bins_x + binA);
        const uint4 a0 = __ldg(pr), a1 = __ldg(pr + 1);
        uint4 b0 = a0, b1 = a1;
        if (binB != binA && binB < F.bins_x) {
            b0 = __ldg(pr + 2);
            b1 = __ldg(pr + 3);
        }
        uint32_t v[4];
        int t_prev = -2;
        double m[6];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int X = X0 + k;
            v[k] = 0u;
            if (X < 0 || X >= oW) continue;
            const bool inB = (X >> 6) != binA;
            const uint4 &r0 = inB ? b0 : a0, &r1 = inB ? b1 : a1;
            const unsigned long long m64 = ((unsigned long long)r0.y << 32) | (unsigned long long)r0.x;
            const int t = pwf_run_id2(r1, (unsigned)__popcll(m64 & ((2ull << (X & 63)) - 1ull)) - 1u);
            if (t < 0) continue;
            if (t != t_prev) {
                pwf_load_matrix(F.inv, t, m);
                t_prev = t;
            }
            const double x = (double)(F.xOff + X);
            v[k] = ldg_or_zero(C.src, pwf_decode<ZERO_OFF>(affine_coord_exact(m[0], x, __dmul_rn(m[2], y), m[4]),
                                                          affine_coord_exact(m[1], x, __dmul_rn(m[3], y), m[5]), F, C.npx_src));
        }
        uint32_t *dst = F.out + ((long long)row * oW + X0);
        if (X0 >= 0 && X0 + 3 < oW) {
            *reinterpret_cast<uint4 *>(dst) = make_uint4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (X0 + k >= 0 && X0 + k < oW) dst[k] = v[k];
        }
    }
}

__global__ void __launch_bounds__(PWF_THREADS, HG_PWF_MINB) pw_warp_fused_kernel(const FusedFrame *frames, int niter)
{
    __shared__ unsigned short s_q[PWF_THREADS / 32][PWF_QCAP];
    const FusedFrame F = frames[blockIdx.y];
    if (F.oW <= 0 || F.oH <= 0) return;
    const int tiles_x = pwf_tiles_x(F.oW);
    const int tile_y = blockIdx.x / tiles_x;
    const int tile_x = blockIdx.x - tile_y * tiles_x;
    const int row0 = tile_y * PWF_GROUP_ROWS * niter;
    if (row0 >= F.oH) return;
    unsigned short *wq = s_q[threadIdx.x >> 5];
    if (F.minSrcX == 0 && F.minSrcY == 0) pwf_body2<true>(F, niter, tile_x, row0, wq);
    else pwf_body2<false>(F, niter, tile_x, row0, wq);
}

// first-generation kernel, kept for A/B runs (HG_PWF_V1=1); needs oW-wide rows of bins == tiles
__global__ void __launch_bounds__(PWF_THREADS, HG_PWF_MINB) pw_warp_fused_v1_kernel(const FusedFrame *frames, int niter)
{
    const FusedFrame F = frames[blockIdx.y];
    if (F.oW <= 0 || F.oH <= 0) return;
    const int tiles_x = pwf_tiles_x(F.oW);
    const int tile_y = blockIdx.x / tiles_x;
    const int tile_x = blockIdx.x - tile_y * tiles_x;
    const int row0 = tile_y * PWF_GROUP_ROWS * niter;
    if (row0 >= F.oH || tile_x >= F.bins_x) return;
    if (F.minSrcX == 0 && F.minSrcY == 0) pwf_body<true>(F, niter, tile_x, row0);
    else pwf_body<false>(F, niter, tile_x, row0);
}

}  // namespace hg
