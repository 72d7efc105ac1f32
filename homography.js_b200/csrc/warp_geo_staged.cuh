// warp_geo_staged.cuh — the opt-in TMA-staged variant of K1/K2 (HG_GEO_STAGED=1): tile classification, the pixel loop
// over a shared-memory copy of the source footprint, and the persistent warp-specialised kernel around them.  Measured
// slower than the direct-gather kernel of warp_geo.cuh on every workload tried (profiles/r01_s2_staged_tma_experiment.md,
// DESIGN.md 3.2b); kept bit-exact and tested as the measured answer to "why not TMA here".  Included at the end of
// warp_geo.cuh (it uses that file's helpers and the direct bodies for the tiles it cannot stage).
#pragma once

namespace hg {

// ------------------------------------------------------------------------------------------------------------
// Source-tile staging through TMA
//
// An output tile (64 pixels x 16*niter rows) of an affine or projective map reads a convex source footprint: the
// image of the tile rectangle is the quadrilateral of its four mapped corners (for the projective map as long as
// the denominator keeps one sign over the tile).  geo_classify maps the corners, takes the bounding box
// (rounded outwards) and sorts the tile into one of four classes:
//   ZERO     the footprint lies entirely outside the image: the tile is transparent, nothing is read;
//   STAGED   footprint inside [0, W-1) x [0, H-1): every pixel is in range and reads a real pixel — the pixel loop
//            needs no bounds test, no flat-index test and no Math.round fix-up (see geo_smem_body);
//   STAGED_CHECK  footprint crosses the image border but stays left of column W-1: out-of-image elements of the
//            box are zero-filled by the TMA unit (= the reference's "index past the end reads undefined -> 0" for
//            row H, Q2); the un-rounded bounds test of H.js:1001 is done per pixel;
//   DIRECT   everything else: footprint reaches column W (where the reference's flat index wraps into the next row,
//            Q2 — a 2-D box cannot express that), box larger than the shared-memory budget (strong minification),
//            no tensor maps (W % 4 != 0 or unaligned source), denominator changing sign or far from 1, huge
//            coordinates.  These tiles run the direct-gather pipeline above (geo_tile_body).
// The staged kernel (warp_inverse_geo_staged_kernel, below) does the classification and the box loads in a producer
// warp that runs ahead of the pixel loops: memory-level parallelism no longer costs registers, which is what
// bounded the direct-gather kernel (ncu: long_scoreboard, 24 warps/SM).
enum { GEO_CLS_DIRECT = 0, GEO_CLS_ZERO = 1, GEO_CLS_STAGED = 2, GEO_CLS_STAGED_CHECK = 3 };

struct GeoTileClass {
    int cls, bx0, by0, pitch, sel, nstrips;
    bool too_big;  // DIRECT only because the box exceeds the shared-memory budget: a shorter tile may still fit
};

// half-width of the zone around a rounding / bounds decision inside which an approximate quotient is not trusted,
// in units of 2^-32 pixel (2^12 -> 2^-20 pixel); the staged projective path folds it into the magic constant
#define HG_NEAR_DELTA_PX (4096.0 / 4294967296.0)

template <int KIND>
__device__ __noinline__ GeoTileClass geo_classify(const GeoFrame &F, const double *mp, int tile_x, int row0, int rows,
                                                  bool has_tm, int box_bytes)
{
    GeoTileClass r;
    r.cls = GEO_CLS_DIRECT;
    r.bx0 = r.by0 = r.pitch = r.sel = r.nstrips = 0;
    r.too_big = false;
    double m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = mp[k];
    // the NOMINAL tile (partial quads and rows past the image included), so every thread's coordinates lie in the box
    const double X0 = (double)(F.xOff + 4 * GEO_TILE_QUADS * tile_x - 3), X1 = (double)(F.xOff + 4 * GEO_TILE_QUADS * tile_x + 4 * GEO_TILE_QUADS - 1);
    const double Y0 = (double)(F.yOff + row0), Y1 = (double)(F.yOff + row0 + rows - 1);
    double minx = 1e300, maxx = -1e300, miny = 1e300, maxy = -1e300, dmin = 1e300, dmax = 0.0;
    bool finite = true, pos = true, neg = true;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const double X = (c & 1) ? X1 : X0, Y = (c & 2) ? Y1 : Y0;
        double sx, sy;
        if (KIND == 0) {
            sx = m[0] * X + m[2] * Y + m[4];
            sy = m[1] * X + m[3] * Y + m[5];
        } else {
            // reciprocal good to 2^-39 relative: far inside the 0.01-pixel margins used below (a zero / NaN / Inf
            // denominator gives non-finite coordinates and the tile goes to the direct path)
            const double dn = m[6] * X + m[7] * Y + 1.0;
            const double rc = rcp_newton1(dn);
            sx = (m[0] * X + m[1] * Y + m[2]) * rc;
            sy = (m[3] * X + m[4] * Y + m[5]) * rc;
            pos = pos && (dn > 0.0);
            neg = neg && (dn < 0.0);
            const double ad = fabs(dn);
            finite = finite && (ad < 1e300);
            dmin = fmin(dmin, ad);
            dmax = fmax(dmax, ad);
        }
        finite = finite && (fabs(sx) < 131072.0) && (fabs(sy) < 131072.0);  // false for NaN / Inf
        minx = fmin(minx, sx);
        maxx = fmax(maxx, sx);
        miny = fmin(miny, sy);
        maxy = fmax(maxy, sy);
    }
    if (!finite) return r;
    if (KIND == 1) {
        // one sign and a moderate range: the tile's image is the convex hull of its corners, the reciprocal needs no
        // exponent guard and the error bound below holds
        if (!(pos || neg) || !(dmin >= 0.015625) || !(dmax <= 64.0)) return r;
    }
    const double W = (double)F.W, H = (double)F.H;
    if (maxx < -0.01 || minx > W + 0.01 || maxy < -0.01 || miny > H + 0.01) {
        r.cls = GEO_CLS_ZERO;
        return r;
    }
    if (!has_tm) return r;
    if (!(maxx <= W - 1.51)) return r;  // Math.round(sx) could reach column W (flat-index wrap, Q2)
    if (KIND == 1) {
        if (m[6] == 0.0 && m[7] == 0.0) return r;  // denominator == 1: the direct path is exact without a reciprocal
        // error of the staged path's quotient against the reference's RN(n / d): the numerators / denominator are
        // formed with a different association (<= 3 ulp of the largest term each) and the reciprocal is good to
        // 2^-39.9 relative; with |q| < 2^17 the bounds below keep the total under 2^-22 pixel << HG_NEAR_DELTA_PX
        const double Xm = fmax(fabs(X0), fabs(X1)), Ym = fmax(fabs(Y0), fabs(Y1));
        const double big = 33554432.0;  // 2^25
        if (!(fabs(m[0]) * Xm + fabs(m[1]) * Ym + fabs(m[2]) < dmin * big)) return r;
        if (!(fabs(m[3]) * Xm + fabs(m[4]) * Ym + fabs(m[5]) < dmin * big)) return r;
        if (!(fabs(m[6]) * Xm + fabs(m[7]) * Ym + 1.0 < dmin * 256.0)) return r;
    }
    // Math.round of every coordinate in [min, max] (+- the 2^-20 the approximate quotient may be off, +- the rounding
    // of this corner arithmetic) lies in [floor(min + 0.49), floor(max + 0.51)].  The TMA unit needs the first
    // column of a box 16-byte aligned in global memory: bx0 is rounded down to a multiple of 4 pixels.
    const int bx0 = ((int)floor(minx + 0.49)) & ~3, bx1 = (int)floor(maxx + 0.51);
    const int by0 = (int)floor(miny + 0.49), by1 = (int)floor(maxy + 0.51);
    const int fw = bx1 - bx0 + 1, fh = by1 - by0 + 1;
    int sel = -1;
#pragma unroll
    for (int i = GEO_NBOX_W - 1; i >= 0; --i)
        if (geo_box_w(i) >= fw) sel = i;
    if (sel < 0) return r;
    const int pitch = geo_box_w(sel);
    const int nstrips = (fh + GEO_BOX_ROWS - 1) / GEO_BOX_ROWS;
    if (nstrips * GEO_BOX_ROWS * pitch * 4 > box_bytes) {
        r.too_big = true;
        return r;
    }
    const bool interior = (minx >= 0.01) && (miny >= 0.01) && (maxy <= H - 1.51);
    r.cls = interior ? GEO_CLS_STAGED : GEO_CLS_STAGED_CHECK;
    r.bx0 = bx0;
    r.by0 = by0;
    r.pitch = pitch;
    r.sel = sel;
    r.nstrips = nstrips;
    return r;
}

// Pixel loop over a staged source tile.  `box` is the shared-memory copy of source rows by0.. / columns bx0.. with
// row pitch `pitch` (out-of-image elements are zero).
//
// Math.round without a fix-up: t = v + (1.5*2^20 + 0.5) puts floor(v + 0.5) = Math.round(v) in the high word.
//   affine      v is formed exactly as the reference does and the magic add rounds DOWN: integers are on the 2^-32
//               grid, so the high word is exact for every v (ties included: v = k + 0.5 gives k + 1, as Math.round).
//   projective  q ~ n * (1/d) with |q - RN(n/d)| < 2^-22; the magic constant also carries +HG_NEAR_DELTA_PX, so the low
//               word of t is frac(q + 0.5) + delta and "q within delta of a rounding boundary" is simply
//               low word < 2*delta.  Two coordinates are tested with ONE multiply: umulhi(lo_x, lo_y) < 2*delta
//               holds whenever either factor is < 2*delta (false positives need both within 2^-9.5 of a boundary;
//               they only cost a trip through the exact path).  Flagged pixels go to the warp queue and are redone
//               by geo_flush_queue with the reference's own arithmetic, reading global memory.
//   CHECK       tiles crossing the image border also need the reference's test on the UNROUNDED coordinate
//               (H.js:1001): floor(v) = round(v) - 1 + (frac(v + 0.5) >= 0.5), then 0 <= floor < W; for the
//               projective map the near zone is widened to every multiple of 0.5 (low word shifted left by one).
template <int KIND, bool CHECK>
__device__ __forceinline__ void geo_smem_body(const GeoFrame &F, const double (&m)[8], int base0, int niter, int s,
                                              int x_first, unsigned mask, uint2 *q, int *qn,
                                              const uint32_t *__restrict__ box, int bx0, int by0, int pitch)
{
    constexpr int R = GEO_ROWS_PER_THREAD;
    const unsigned W = (unsigned)F.W, H = (unsigned)F.H;
    const int oH = F.oH;
    const long long row_pitch = (long long)F.oW;
    const double MG = (KIND == 0) ? (HG_MAGIC + 0.5) : (HG_MAGIC + 0.5 + HG_NEAR_DELTA_PX);
    // shared-memory byte address of source pixel (rx, ry) = hi(ty) * 4*pitch + (hi(tx) * 4 + kaddr): the magic
    // exponent bits and the box origin are folded into one constant (32-bit wrap-around arithmetic)
    const unsigned pitch4 = 4u * (unsigned)pitch;
    const unsigned kaddr = smem_u32(box) - pitch4 * (unsigned)(HG_HI_ZERO + by0) - 4u * (unsigned)(HG_HI_ZERO + bx0);
    double xs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) xs[k] = (double)(F.xOff + x_first + k);

#pragma unroll 1
    for (int it = 0; it < niter; ++it) {
        const int base = base0 + it * GEO_GROUP_ROWS;
        if (base >= oH) break;
        uint32_t px[R][4];
        unsigned redo_bits = 0u;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const double y = (double)(F.yOff + base + s * j);
            double r0, r1, r2 = 0.0;
            if (KIND == 0) {
                r0 = __dmul_rn(m[2], y);
                r1 = __dmul_rn(m[3], y);
            } else {
                r0 = __fma_rn(m[1], y, m[2]);
                r1 = __fma_rn(m[4], y, m[5]);
                r2 = __fma_rn(m[7], y, 1.0);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                double tx, ty;
                if (KIND == 0) {
                    tx = __dadd_rd(affine_coord_exact(m[0], xs[k], r0, m[4]), MG);
                    ty = __dadd_rd(affine_coord_exact(m[1], xs[k], r1, m[5]), MG);
                } else {
                    const double rc = rcp_newton1(__fma_rn(m[6], xs[k], r2));
                    tx = __fma_rn(__fma_rn(m[0], xs[k], r0), rc, MG);
                    ty = __fma_rn(__fma_rn(m[3], xs[k], r1), rc, MG);
                    const unsigned lx = (unsigned)__double2loint(tx), ly = (unsigned)__double2loint(ty);
                    const bool again = CHECK ? (__umulhi(lx << 1, ly << 1) < 4u * HG_NEAR_DELTA)
                                             : (__umulhi(lx, ly) < 2u * HG_NEAR_DELTA);
                    redo_bits |= again ? (1u << (4 * j + k)) : 0u;
                }
                const unsigned hx = (unsigned)__double2hiint(tx), hy = (unsigned)__double2hiint(ty);
                uint32_t v = lds_u32(hy * pitch4 + (hx * 4u + kaddr));
                if (CHECK) {
                    const unsigned ux = hx - (unsigned)(HG_HI_ZERO + 1) + ((unsigned)__double2loint(tx) >> 31);
                    const unsigned uy = hy - (unsigned)(HG_HI_ZERO + 1) + ((unsigned)__double2loint(ty) >> 31);
                    v = ((ux < W) & (uy < H)) ? v : 0u;
                }
                px[j][k] = v;
            }
        }
        if (KIND == 1 && redo_bits) {
            // one entry per thread and row group; the queue is emptied after every tile and holds 32 * niter entries
            // (GEO_QCAP >= 32 * niter is checked on the host), so it cannot overflow here
            q[atomicAdd(qn, 1)] = make_uint2((unsigned)(x_first + 4), (redo_bits << 17) | (unsigned)base);
        }
        uint32_t *dst = F.out + ((long long)base * row_pitch + x_first);
#pragma unroll
        for (int j = 0; j < R; ++j) {
            if (base + s * j < oH) {
                if (mask == 0xFu) {
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(px[j][0], px[j][1], px[j][2], px[j][3]);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (mask & (1u << k)) dst[k] = px[j][k];
                }
            }
            dst += (long long)s * row_pitch;
        }
    }
}

// a tile whose footprint misses the image: transparent
__device__ __forceinline__ void geo_zero_body(const GeoFrame &F, int base0, int niter, int s, int x_first, unsigned mask)
{
    const long long row_pitch = (long long)F.oW;
    for (int it = 0; it < niter; ++it) {
        const int base = base0 + it * GEO_GROUP_ROWS;
        if (base >= F.oH) break;
        uint32_t *dst = F.out + ((long long)base * row_pitch + x_first);
#pragma unroll
        for (int j = 0; j < GEO_ROWS_PER_THREAD; ++j) {
            if (base + s * j < F.oH) {
                if (mask == 0xFu) {
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(0u, 0u, 0u, 0u);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (mask & (1u << k)) dst[k] = 0u;
                }
            }
            dst += (long long)s * row_pitch;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Staged kernel: persistent, warp-specialised.
//
// CTA = 4 consumer warps (the 16 x 8 thread tile of the direct kernel) + 1 producer warp, several CTAs per SM, each
// walking the tile sequence t = blockIdx.x, blockIdx.x + gridDim.x, ... over all frames of the launch.  Shared memory
// holds a ring of `stages` slots (header + staged source box) with a full / empty mbarrier pair each:
//   producer  lane l classifies tile k0 + l of the next 32 tiles (corner maps, bounding box, class — all lanes in
//             parallel; a tile whose box exceeds the slot is cut into two half-height entries and classified again),
//             then the lanes take turns in tile order: wait for the slot to be empty, write the header (frame, matrix,
//             class, box origin, rows), and either arm the full barrier with the box's byte count and issue the TMA
//             box loads, or (zero / direct entries) just arrive on it.  An END entry closes the sequence;
//   consumers wait on the full barrier, run the class's pixel loop (geo_smem_body / geo_zero_body / geo_tile_body),
//             resolve the entry's queued exact pixels, and one lane per warp arrives on the empty barrier.
// The producer runs `stages` entries ahead, so classification and the HBM latency of the box loads are off the
// consumers' critical path and the bytes in flight per SM (what bounds a gather kernel) are set by the ring depth,
// not by registers.
constexpr int GEO_MAX_STAGES = 6;
constexpr int GEO_STAGED_THREADS = GEO_THREADS + 32;
enum { GEO_CLS_END = -1 };

struct GeoStageHdr {
    GeoFrame F;
    double m[8];
    int cls, bx0, by0, pitch, tile_x, row0, niter, pad;
};
constexpr int GEO_HDR_BYTES = 256;  // sizeof(GeoStageHdr) rounded up so the box behind it stays 128-byte aligned
static_assert(sizeof(GeoStageHdr) <= GEO_HDR_BYTES, "stage header does not fit");

template <int KIND>
__device__ __forceinline__ void geo_load_frame(const GeoParams &P, int frame, GeoFrame &F, double (&m)[8])
{
    F = P.many ? P.many[frame] : P.one;
    if (P.mats_dev) {
        if (KIND == 0) {
            const float *mf = (const float *)P.mats_dev + 6 * (size_t)frame;
#pragma unroll
            for (int k = 0; k < 6; ++k) m[k] = (double)__ldg(mf + k);
            m[6] = m[7] = 0.0;
        } else {
            const double *md = (const double *)P.mats_dev + 8 * (size_t)frame;
#pragma unroll
            for (int k = 0; k < 8; ++k) m[k] = __ldg(md + k);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = P.mat_val[k];
    }
}

// the rare paths of the staged kernel live behind calls so their registers do not count against the staged pixel loop
template <int KIND>
__device__ __noinline__ void geo_direct_tile(const GeoStageHdr *hp, int base, int s, int x_first, unsigned mask, uint2 *q, int *qn)
{
    const GeoFrame F = hp->F;
    const int niter = hp->niter;
    double m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = hp->m[i];
    if (KIND == 0) {
        geo_tile_body<0, 0>(F, m, base, niter, s, x_first, mask, q, qn);
    } else {
        // tile-uniform mode from the matrix and the frame window
        const double ax = fmax(fabs((double)F.xOff), fabs((double)F.xOff + (double)F.oW));
        const double ay = fmax(fabs((double)F.yOff), fabs((double)F.yOff + (double)F.oH));
        const double spread = fabs(m[6]) * ax + fabs(m[7]) * ay;  // |h6 x + h7 y| <= spread (+ rounding)
        if (m[6] == 0.0 && m[7] == 0.0) geo_tile_body<1, 2>(F, m, base, niter, s, x_first, mask, q, qn);
        else if (spread < 0.75) geo_tile_body<1, 1>(F, m, base, niter, s, x_first, mask, q, qn);
        else geo_tile_body<1, 0>(F, m, base, niter, s, x_first, mask, q, qn);
    }
}

__device__ __noinline__ void geo_flush_tile(const GeoStageHdr *hp, const uint2 *q, int n, int s, int lane)
{
    const GeoFrame F = hp->F;
    double m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = hp->m[i];
    geo_flush_queue<false>(F, m, q, n, s, lane);
}

// one ring entry: wait for the slot, publish the header, start the box loads (or just signal)
__device__ __forceinline__ void geo_emit(const GeoStageHdr &h, const CUtensorMap *tm, int sel, int nstrips,
                                         unsigned char *slot_base, uint32_t full_bar, uint32_t empty_bar, unsigned use,
                                         int debug = 0)
{
    if (debug & 1)
        printf("hgwarp: block %d emit cls %d tile_x %d row0 %d niter %d box (%d,%d) pitch %d sel %d strips %d use %u tm %p\n",
               (int)blockIdx.x, h.cls, h.tile_x, h.row0, h.niter, h.bx0, h.by0, h.pitch, sel, nstrips, use, (const void *)tm);
    if (use > 0) mbar_wait(empty_bar, (use - 1) & 1, 1);
    *reinterpret_cast<GeoStageHdr *>(slot_base) = h;
    if (h.cls >= GEO_CLS_STAGED) {
        // nstrips counts 8-row strips; groups of four go out as one 32-row box
        const unsigned strip_bytes = (unsigned)(GEO_BOX_ROWS * h.pitch * 4);
        mbar_arrive_expect_tx(full_bar, strip_bytes * (unsigned)nstrips);
        uint32_t dst = smem_u32(slot_base + GEO_HDR_BYTES);
        int y = h.by0, left = nstrips;
        constexpr int TALL = GEO_BOX_ROWS_TALL / GEO_BOX_ROWS;
        for (; left >= TALL && !(debug & 2); left -= TALL) {
            tma_load_2d(dst, tm + GEO_NBOX_W + sel, h.bx0, y, full_bar);
            dst += TALL * strip_bytes;
            y += GEO_BOX_ROWS_TALL;
        }
        for (; left > 0; --left) {
            tma_load_2d(dst, tm + sel, h.bx0, y, full_bar);
            dst += strip_bytes;
            y += GEO_BOX_ROWS;
        }
    } else {
        mbar_arrive(full_bar);
    }
}

// producer warp of the staged kernel (see the kernel's header comment)
template <int KIND>
__device__ __noinline__ void geo_producer(const GeoParams &P, unsigned char *s_ring, unsigned long long *s_full,
                                          unsigned long long *s_empty)
{
    const int lane_id = threadIdx.x & 31;
    const unsigned S = (unsigned)P.stages;
    const unsigned slot_bytes = (unsigned)(GEO_HDR_BYTES + P.box_bytes);
    const unsigned tiles_x_max = (unsigned)P.tiles_x, tpf = (unsigned)(P.tiles_x * P.tiles_y);
    const unsigned total = tpf * (unsigned)P.n_frames;  // < 2^31: the host splits larger batches
    const int niter = P.niter, rows = GEO_GROUP_ROWS * niter;
    unsigned slot = 0, use = 0;  // ring position of the next entry
    for (unsigned k0 = 0;; k0 += 32) {
        if (blockIdx.x + k0 * gridDim.x >= total) break;
        const unsigned t = blockIdx.x + (k0 + lane_id) * gridDim.x;
        GeoStageHdr h;
        h.cls = GEO_CLS_DIRECT;
        h.bx0 = h.by0 = h.pitch = h.tile_x = h.row0 = h.pad = 0;
        h.niter = niter;
        // second entry of a tile cut in two (same frame, tile column and pitch family)
        int n_ent = 0, cls2 = GEO_CLS_DIRECT, bx2 = 0, by2 = 0, pitch2 = 0, sel2 = 0, nstrips2 = 0;
        int sel = 0, nstrips = 0;
        const CUtensorMap *tm = nullptr;
        if (t < total) {
            const unsigned frame = t / tpf, r = t - frame * tpf;
            const unsigned tile_y = r / tiles_x_max;
            h.tile_x = (int)(r - tile_y * tiles_x_max);
            h.row0 = (int)tile_y * rows;
            geo_load_frame<KIND>(P, (int)frame, h.F, h.m);
            if (h.tile_x < geo_tiles_x(h.F.oW) && h.row0 < h.F.oH) {  // else: outside this (smaller) frame of the batch
                n_ent = 1;
                tm = P.many ? h.F.tm : (P.has_tm ? P.tm_val : nullptr);
                GeoTileClass tc = geo_classify<KIND>(h.F, h.m, h.tile_x, h.row0, rows, tm != nullptr, P.box_bytes);
                if (tc.too_big && (niter & 1) == 0) {
                    const int half = rows / 2;
                    h.niter = niter / 2;
                    tc = geo_classify<KIND>(h.F, h.m, h.tile_x, h.row0, half, true, P.box_bytes);
                    if (h.row0 + half < h.F.oH) {
                        n_ent = 2;
                        const GeoTileClass t2 = geo_classify<KIND>(h.F, h.m, h.tile_x, h.row0 + half, half, true, P.box_bytes);
                        cls2 = t2.cls;
                        bx2 = t2.bx0;
                        by2 = t2.by0;
                        pitch2 = t2.pitch;
                        sel2 = t2.sel;
                        nstrips2 = t2.nstrips;
                    }
                }
                h.cls = tc.cls;
                h.bx0 = tc.bx0;
                h.by0 = tc.by0;
                h.pitch = tc.pitch;
                sel = tc.sel;
                nstrips = tc.nstrips;
            }
        }
        for (int i = 0; i < 32; ++i) {
            if (blockIdx.x + (k0 + i) * gridDim.x >= total) break;  // warp-uniform
            const int ne = __shfl_sync(0xFFFFFFFFu, n_ent, i);
            for (int e = 0; e < ne; ++e) {
                if (lane_id == i) {
                    if (e == 1) {
                        h.row0 += GEO_GROUP_ROWS * h.niter;
                        h.cls = cls2;
                        h.bx0 = bx2;
                        h.by0 = by2;
                        h.pitch = pitch2;
                        sel = sel2;
                        nstrips = nstrips2;
                    }
                    geo_emit(h, tm, sel, nstrips, s_ring + (size_t)slot * slot_bytes, smem_u32(&s_full[slot]),
                             smem_u32(&s_empty[slot]), use, P.debug);
                }
                if (++slot == S) {
                    slot = 0;
                    ++use;
                }
            }
            __syncwarp();
        }
    }
    if (lane_id == 0) {
        GeoStageHdr h;
        h.cls = GEO_CLS_END;
        geo_emit(h, nullptr, 0, 0, s_ring + (size_t)slot * slot_bytes, smem_u32(&s_full[slot]), smem_u32(&s_empty[slot]), use,
                 P.debug);
    }
}

template <int KIND>
__global__ void __launch_bounds__(GEO_STAGED_THREADS, HG_GEO_STAGED_MINB) warp_inverse_geo_staged_kernel(const __grid_constant__ GeoParams P)
{
    extern __shared__ __align__(128) unsigned char s_ring[];  // stages x (header + box)
    __shared__ uint2 s_q[KIND == 1 ? GEO_THREADS / 32 : 1][KIND == 1 ? GEO_QCAP : 1];
    __shared__ int s_qn[GEO_THREADS / 32];
    __shared__ __align__(8) unsigned long long s_full[GEO_MAX_STAGES], s_empty[GEO_MAX_STAGES];

    const int warp_id = threadIdx.x >> 5, lane_id = threadIdx.x & 31;
    const unsigned S = (unsigned)P.stages;
    const unsigned slot_bytes = (unsigned)(GEO_HDR_BYTES + P.box_bytes);
    if (threadIdx.x == 0) {
#pragma unroll 1
        for (unsigned i = 0; i < S; ++i) {
            mbar_init(smem_u32(&s_full[i]), 1);
            mbar_init(smem_u32(&s_empty[i]), GEO_THREADS / 32);
        }
        fence_barrier_init();
    }
    if (lane_id == 0 && warp_id < GEO_THREADS / 32) s_qn[warp_id] = 0;
    __syncthreads();

    if (warp_id == GEO_THREADS / 32) {
        geo_producer<KIND>(P, s_ring, s_full, s_empty);
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    uint2 *q = s_q[KIND == 1 ? warp_id : 0];
    int *qn = &s_qn[warp_id];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    unsigned slot = 0, use = 0;
    for (;;) {
        mbar_wait(smem_u32(&s_full[slot]), use & 1, 2);
        const unsigned char *base_p = s_ring + (size_t)slot * slot_bytes;
        const GeoStageHdr *hp = reinterpret_cast<const GeoStageHdr *>(base_p);
        const int cls = hp->cls;
        if ((P.debug & 1) && lane_id == 0) printf("hgwarp: block %d warp %d got cls %d slot %u use %u\n", (int)blockIdx.x, warp_id, cls, slot, use);
        if (cls == GEO_CLS_END) break;
        {
            const GeoFrame F = hp->F;
            const int oW = F.oW, oH = F.oH, niter = hp->niter;
            // rows with equal flat alignment repeat with period s = 4 / gcd(oW mod 4, 4)
            const int sl = (oW & 3) == 0 ? 0 : ((oW & 1) ? 2 : 1);  // log2(s)
            const int s = 1 << sl;
            const int base = hp->row0 + (ty >> sl) * (s * GEO_ROWS_PER_THREAD) + (ty & (s - 1));
            const int shift = (int)(((unsigned)base * (unsigned)oW) & 3u);
            const int x_first = 4 * (hp->tile_x * GEO_TILE_QUADS + tx) - shift;
            const bool active = (base < oH) && (x_first < oW);
            unsigned mask = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (x_first + i >= 0 && x_first + i < oW) mask |= 1u << i;
            if (active) {
                if (cls >= GEO_CLS_STAGED) {
                    double m[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) m[i] = hp->m[i];
                    const uint32_t *box = reinterpret_cast<const uint32_t *>(base_p + GEO_HDR_BYTES);
                    if (cls == GEO_CLS_STAGED) geo_smem_body<KIND, false>(F, m, base, niter, s, x_first, mask, q, qn, box, hp->bx0, hp->by0, hp->pitch);
                    else geo_smem_body<KIND, true>(F, m, base, niter, s, x_first, mask, q, qn, box, hp->bx0, hp->by0, hp->pitch);
                } else if (cls == GEO_CLS_ZERO) {
                    geo_zero_body(F, base, niter, s, x_first, mask);
                } else {
                    geo_direct_tile<KIND>(hp, base, s, x_first, mask, q, qn);
                }
            }
            if (KIND == 1) {
                // resolve the entry's queued pixels (the __syncwarp also orders the provisional stores before the
                // corrected ones); the header is still needed, so this comes before the slot is released
                __syncwarp();
                const int n = min(*qn, GEO_QCAP);
                if (n > 0) {
                    geo_flush_tile(hp, q, n, s, lane_id);
                    __syncwarp();
                    if (lane_id == 0) *qn = 0;
                }
            }
            // the slot (header and box) is no longer needed by this warp
            __syncwarp();
            if (lane_id == 0) mbar_arrive(smem_u32(&s_empty[slot]));
        }
        if (++slot == S) {
            slot = 0;
            ++use;
        }
    }
}

}  // namespace hg
