// warp_geo.cuh — K1/K2: the pixel loop of _inverseGeometricWarp (H.js:997-1011) as an HBM-bound
// gather kernel.
//
//   for every output pixel (x,y) in [xOff,xOff+oW) x [yOff,yOff+oH):
//       (sx,sy) = T^-1(x,y)                      affine: H.js:1382, projective: H.js:1401
//       if 0 <= sx < W and 0 <= sy < H:          H.js:1001 (test on the UNROUNDED coordinate)
//           out[x,y] = src_flat[round(sy)*W + round(sx)]   H.js:1005 (flat index, Math.round;
//                                                 index past the end reads `undefined` -> 0)
//
// Data layout.  RGBA8 pixels are 32-bit words; the output is dense (row pitch = oW pixels, the layout of
// ImageData), so a row starts at flat pixel yy*oW whose 16-byte alignment is (yy*oW) mod 4.  The kernel
// walks each row in FLAT-aligned quads: quad i of row yy covers x = 4i - a(yy) .. +3 with a(yy) = (yy*oW)&3,
// so every interior quad is ONE aligned 128-bit store for any oW (the first/last quad of a row may be
// partial and is stored pixel by pixel).  a(yy) repeats with period s = 4/gcd(oW mod 4, 4), so a thread
// that owns rows yy, yy+s, yy+2s, ... sees the same x for all of them.
//
// Work decomposition.  CTA = 128 threads = 16 (x) x 8 (y); thread (tx,ty) owns one quad column and R = 2 rows per row
// group, `niter` row groups 16 rows apart -> CTA tile = 64 pixels x 16*niter rows (niter is picked on the host so the
// grid still fills the 148 SMs >= 16 times).  A warp (16 lanes x 2 thread rows) stores 256 contiguous bytes per row.
//
// Pixel loops in this file (all bit-exact with the reference's unfused doubles):
//   geo_fast_body   the default: doubled coordinates (n = floor(2v + 1) from ONE fma; Math.round = n >> 1, the bounds
//                   test of H.js:1001 = one unsigned compare), two-stage software pipeline with predicated gathers,
//                   warp-uniform "end pixels inside" shortcut, L2 prefetch one row group ahead.  Affine: exact, no
//                   fix-up.  Projective (denominator of one sign and moderate size over the window, geo_fast_mode):
//                   MUFU.RCP64H + one Newton step, quotients within 2^-19 of a decision boundary are queued per warp
//                   and resolved exactly after the loop (geo_flush_queue<true>, no division: quotient_at_least).
//   geo_tile_body   first-generation body, kept for the projective frames geo_fast_mode rejects (horizon inside the
//                   window, denominator == 1, extreme magnitudes): reference-exact numerators, exponent guard, IEEE
//                   divisions in the exact path; three-stage pipeline with a sentinel flat index.
//   geo_smem_body   pixel loop of the opt-in TMA-staged kernel (warp_inverse_geo_staged_kernel, HG_GEO_STAGED=1).
#pragma once
#include "jsnum.cuh"
#include "tma.cuh"

namespace hg {

// Source-tile staging (TMA).  One tensor map per box width and height: the source footprint of an output tile is
// staged as a stack of boxes geo_box_w(sel) pixels wide — first the 32-row ones, then 8-row ones for the rest — so
// shared memory holds rows by0.. with row pitch geo_box_w(sel) and a staged pixel is addressed as
// (ry - by0) * pitch + (rx - bx0).
constexpr int GEO_NBOX_W = 8;               // box widths
constexpr int GEO_NBOX = 2 * GEO_NBOX_W;    // x two box heights: maps [0,8) are GEO_BOX_ROWS tall, [8,16) GEO_BOX_ROWS_TALL
constexpr int GEO_BOX_ROWS = 8;
constexpr int GEO_BOX_ROWS_TALL = 32;
__host__ __device__ inline int geo_box_w(int i)
{
    return i == 0 ? 72 : i == 1 ? 80 : i == 2 ? 96 : i == 3 ? 112 : i == 4 ? 128 : i == 5 ? 160 : i == 6 ? 192 : 256;
}

struct GeoFrame {
    const uint32_t *src;
    uint32_t *out;
    int W, H;
    int xOff, yOff, oW, oH;
    const CUtensorMap *tm;   // device array of GEO_NBOX tensor maps over src, or nullptr (no staging)
};

struct GeoParams {
    GeoFrame one;            // used when many == nullptr
    const GeoFrame *many;    // device array, indexed by blockIdx.y
    const void *mats_dev;    // device matrices (float[6] | double[8] per frame) or nullptr
    int niter;               // row groups per CTA
    int ltx;                 // log2(quads per CTA row): 4 = the default 64 x 16 pixel row group; 1 = "tall" 8 x 128 (rotated maps)
    int has_tm;              // tm_val holds the tensor maps of one.src (single-frame launches)
    int box_bytes;           // shared memory for one staged source box (staged kernel)
    int stages;              // ring depth of the staged kernel
    int tiles_x, tiles_y;    // tile grid of the largest frame (staged kernel: tile index -> frame, tile)
    int n_frames;
    int debug;               // HG_GEO_DEBUG: the staged kernel's producer prints every ring entry
    double mat_val[8];       // matrix by value when mats_dev == nullptr (floats widened to double)
    CUtensorMap tm_val[GEO_NBOX];
};

constexpr int GEO_TILE_QUADS = 16;  // quads per CTA row (64 pixels)
constexpr int GEO_TY = 8;           // thread rows per CTA  -> 128 threads
#ifndef HG_GEO_R
#define HG_GEO_R 2
#endif
#ifndef HG_GEO_MINB
#define HG_GEO_MINB 6  // CTAs per SM the affine kernel is compiled for (80 registers)
#endif
#ifndef HG_GEO_MINB_PROJ
// projective: 7 CTAs per SM (72 registers, 8 bytes of spill outside the loop).  Measured on config 2 / generic points after the
// last changes to the loop: 5 CTAs (96 registers) 0.782 / 0.798 of the HBM bound, 6 (80) 0.778 / 0.792, 7 (72) 0.796 / 0.806,
// 8 (64, 128 bytes of spill loads) 0.764.  (Round 1's loop preferred 5 over 6 by 3 %.)
#define HG_GEO_MINB_PROJ 7
#endif
#ifndef HG_GEO_MINB_TALL
// affine maps near a quarter turn (tall thread layout, GeoParams::ltx == 1): their gathers walk down source columns and are
// bound by latency, not by issue slots — 8 CTAs per SM (64 registers, no spills): 0.766 -> 0.819 of the HBM bound (7: 0.804);
// the same 8 CTAs cost a translation-like map 3.7 % (0.946 -> 0.911), hence a kernel instance of its own
#define HG_GEO_MINB_TALL 8
#endif
#ifndef HG_GEO_STAGED_MINB
#define HG_GEO_STAGED_MINB 5
#endif
constexpr int GEO_ROWS_PER_THREAD = HG_GEO_R;
constexpr int GEO_THREADS = GEO_TILE_QUADS * GEO_TY;
constexpr int GEO_GROUP_ROWS = GEO_TY * GEO_ROWS_PER_THREAD;  // rows one CTA covers per iteration
constexpr int GEO_QCAP = 96;  // queued (thread, row group) entries per warp awaiting exact resolution

// host + device: CTAs needed for one frame.  Rows whose flat start is not 16-byte aligned begin with a partial
// quad, so a row has at most (oW + 3 + 3) / 4 quads; when oW % 4 == 0 every row is aligned.
__host__ __device__ inline int geo_tiles_x(int oW, int tile_quads = GEO_TILE_QUADS)
{
    const int quads = (oW & 3) == 0 ? oW / 4 : (oW + 6) / 4;
    return (quads + tile_quads - 1) / tile_quads;
}
__host__ __device__ inline int geo_tiles_y(int oH, int niter, int group_rows = GEO_GROUP_ROWS)
{
    return (oH + group_rows * niter - 1) / (group_rows * niter);
}
// The 128 threads of a CTA form 2^ltx quad columns x (128 >> ltx) thread rows.  ltx = 4 is the default (a warp stores
// 256 contiguous bytes per row).  ltx = 1 ("tall") is for maps that turn the image by about a quarter turn: there a
// step in output x is a step in source y, so lanes laid out along output ROWS read along a source row — coalesced
// gathers again — and a warp still stores whole 32-byte sectors (two adjacent quads per row).
__host__ __device__ inline int geo_group_rows(int ltx) { return (GEO_THREADS >> ltx) * GEO_ROWS_PER_THREAD; }

// hi word of (v + 1.5*2^20) minus the hi word of (0 + 1.5*2^20): equals floor(v) for -2^19 <= v < 2^19 and is
// >= 2^19 (as unsigned) for everything else incl. NaN / Inf, so `(unsigned)u < W` is the complete test
// "v is finite and 0 <= v < W".
#define HG_HI_ZERO 0x41380000

#define HG_NEAR_DELTA 4096u  // 2^12 * 2^-32 = 2^-20 absolute

// exact quotient path: t = RD(RN(n / d) + magic) — the reference's own arithmetic
__device__ __noinline__ double exact_quotient_magic(double n, double d)
{
    return __dadd_rd(__ddiv_rn(n, d), HG_MAGIC);
}

__device__ __forceinline__ double rcp_newton1(double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    const double e = __fma_rn(-d, r, 1.0);
    return __fma_rn(r, e, r);
}

__device__ __forceinline__ bool near_half_multiple(unsigned frac)
{
    // frac within DELTA of 0, 2^31 or 2^32  <=>  ((frac + DELTA) mod 2^31) < 2*DELTA
    return ((frac + HG_NEAR_DELTA) << 1) < 4u * HG_NEAR_DELTA;
}

#define HG_OUTSIDE 0xFFFFFFFFu  // flat-index sentinel: the pixel reads nothing and becomes transparent

// 32-bit gather that returns 0 without touching memory when idx is the OUTSIDE sentinel
__device__ __forceinline__ uint32_t ldg_or_zero(const uint32_t *__restrict__ src, unsigned idx)
{
    uint32_t v;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0xffffffff;\n\tmov.b32 %0, 0;\n\t@q ld.global.nc.b32 %0, [%1+0];\n\t}"
        : "=r"(v)
        : "l"(src + idx), "r"(idx));
    return v;
}

// decode t = coordinate + magic into the flat source index (or HG_OUTSIDE); see Q1/Q2 in the header comment
__device__ __forceinline__ unsigned decode_flat(double tx, double ty, unsigned W, unsigned H, unsigned npx_src)
{
    const unsigned ux = (unsigned)(__double2hiint(tx) - HG_HI_ZERO);
    const unsigned uy = (unsigned)(__double2hiint(ty) - HG_HI_ZERO);
    // Math.round = floor + (frac >= 0.5); the flat index may run past the row end (reads the next row, Q2)
    const unsigned rx = ux + ((unsigned)__double2loint(tx) >> 31);
    const unsigned ry = uy + ((unsigned)__double2loint(ty) >> 31);
    const unsigned flat = ry * W + rx;
    return ((ux < W) & (uy < H) & (flat < npx_src)) ? flat : HG_OUTSIDE;
}

// MODE (projective only): 0 general + denominator exponent guard, 1 general (denominator proven in
// (0.25,1.75)), 2 denominator == 1 exactly.  (A fourth mode — one reciprocal per column when h7 == 0 — was built and
// measured: its extra live registers cost more occupancy than the 5 FP64 ops per pixel it saved, so it was dropped.)
//
// Per thread: one quad column, GEO_ROWS_PER_THREAD rows per iteration, `niter` iterations (GEO_GROUP_ROWS rows
// apart), run as a THREE-STAGE SOFTWARE PIPELINE so that no instruction ever waits on the stage before it:
//   iteration i:   C  store the pixels gathered for group i-2          (their loads were issued one iteration ago)
//                  B  fix up flagged pixels of group i-1 (rare), then issue its 32-bit gathers
//                  A  arithmetic of group i: coordinates -> flat source index (+ "redo" bit)
// Stage A is straight-line code over the thread's pixels, so the FP64 dependency chains of different pixels
// interleave; its results are first consumed in the NEXT iteration, and memory latency is covered by a whole
// iteration of arithmetic inside the same warp.
template <int KIND, int MODE>
__device__ __forceinline__ void geo_tile_body(const GeoFrame &F, const double (&m)[8], int base0, int niter, int s,
                                              int x_first, unsigned mask, uint2 *q, int *qn, int group_rows = GEO_GROUP_ROWS)
{
    static_assert(4 * GEO_ROWS_PER_THREAD + 17 <= 32, "queue entry packs the redo bits above a 17-bit row");
    constexpr int R = GEO_ROWS_PER_THREAD;
    const uint32_t *__restrict__ src = F.src;
    const unsigned W = (unsigned)F.W, H = (unsigned)F.H;
    const unsigned npx_src = W * H;  // < 2^31 (checked on the host)
    const int oH = F.oH;

    // x-only terms, shared by every row this thread touches
    double xs[4], ax0[4], ax1[4], ax2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        xs[k] = (double)(F.xOff + x_first + k);
        if (KIND == 1) {
            ax0[k] = __dmul_rn(m[0], xs[k]);
            ax1[k] = __dmul_rn(m[3], xs[k]);
            ax2[k] = (MODE == 2) ? 0.0 : __dmul_rn(m[6], xs[k]);
        }
    }

    uint32_t px[R][4];       // stage C operands (gathers in flight)
    unsigned idx[R][4];      // stage B operands (flat indices of the previous group)
    unsigned redo_bits = 0;  // pixels of the previous group whose quotient must be resolved exactly
    int qpos = 0;            // queue slot reserved for them
    int base_px = -1, base_idx = -1;
    const long long row_pitch = (long long)F.oW;

#pragma unroll 1
    for (int it = 0;; ++it) {
        // ---- C: store group it-2
        if (base_px >= 0) {
            uint32_t *dst = F.out + ((long long)base_px * row_pitch + x_first);
#pragma unroll
            for (int j = 0; j < R; ++j) {
                if (base_px + s * j < oH) {
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
            base_px = -1;
        }
        // ---- B: gathers of group it-1
        if (base_idx >= 0) {
            if (KIND == 1 && MODE != 2 && redo_bits) {
                // Flagged pixels (quotient within 2^-20 of a decision boundary) are NOT resolved here, where one
                // lane would run the IEEE divide while 31 wait.  The thread appends ONE entry (quad column, row
                // group, flag bits) to its warp's queue — the slot was reserved one iteration ago, so the atomic's
                // latency is hidden — and the queue is resolved after the loop (geo_flush_queue), overwriting the
                // provisional pixels stored below.  Only when the queue is full is a thread resolved in place.
                if (qpos < GEO_QCAP) {
                    q[qpos] = make_uint2((unsigned)(x_first + 4), (redo_bits << 17) | (unsigned)base_idx);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        unsigned rows = (redo_bits >> k) & 0x11111111u;
                        while (rows) {
                            const int j = (__ffs((int)rows) - 1) >> 2;
                            rows &= rows - 1;
                            const double y = (double)(F.yOff + base_idx + s * j);
                            const double nx = __dadd_rn(__dadd_rn(ax0[k], __dmul_rn(m[1], y)), m[2]);
                            const double ny = __dadd_rn(__dadd_rn(ax1[k], __dmul_rn(m[4], y)), m[5]);
                            const double dn = __dadd_rn(__dadd_rn(ax2[k], __dmul_rn(m[7], y)), 1.0);
                            const unsigned f = decode_flat(exact_quotient_magic(nx, dn), exact_quotient_magic(ny, dn), W,
                                                           H, npx_src);
#pragma unroll
                            for (int jj = 0; jj < R; ++jj)
                                if (jj == j) idx[jj][k] = f;  // select, no dynamic register indexing
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < R; ++j)
#pragma unroll
                for (int k = 0; k < 4; ++k) px[j][k] = ldg_or_zero(src, idx[j][k]);
            base_px = base_idx;
            base_idx = -1;
        }
        // ---- A: arithmetic of group it
        const int base = base0 + it * group_rows;
        if (it < niter && base < oH) {
            redo_bits = 0u;
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const double y = (double)(F.yOff + base + s * j);
                double r0, r1, r2 = 0.0;
                if (KIND == 0) {
                    r0 = __dmul_rn(m[2], y);
                    r1 = __dmul_rn(m[3], y);
                } else {
                    r0 = __dmul_rn(m[1], y);
                    r1 = __dmul_rn(m[4], y);
                    if (MODE != 2) r2 = __dmul_rn(m[7], y);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    double tx, ty;
                    if (KIND == 0) {
                        tx = __dadd_rd(affine_coord_exact(m[0], xs[k], r0, m[4]), HG_MAGIC);
                        ty = __dadd_rd(affine_coord_exact(m[1], xs[k], r1, m[5]), HG_MAGIC);
                    } else {
                        const double nx = __dadd_rn(__dadd_rn(ax0[k], r0), m[2]);
                        const double ny = __dadd_rn(__dadd_rn(ax1[k], r1), m[5]);
                        if (MODE == 2) {
                            // denominator is exactly 1: n / 1 = n, no approximation anywhere
                            tx = __dadd_rd(nx, HG_MAGIC);
                            ty = __dadd_rd(ny, HG_MAGIC);
                        } else {
                            const double dn = __dadd_rn(__dadd_rn(ax2[k], r2), 1.0);
                            const double rc = rcp_newton1(dn);
                            tx = __fma_rn(nx, rc, HG_MAGIC);
                            ty = __fma_rn(ny, rc, HG_MAGIC);
                            bool again = near_half_multiple((unsigned)__double2loint(tx)) |
                                         near_half_multiple((unsigned)__double2loint(ty));
                            if (MODE == 0) {
                                // |dn| outside [2^-500, 2^500], or 0 / Inf / NaN: reciprocal not trusted
                                const unsigned de = ((unsigned)__double2hiint(dn) & 0x7FF00000u) - (523u << 20);
                                again |= de > (1000u << 20);
                            }
                            redo_bits |= again ? (1u << (4 * j + k)) : 0u;
                        }
                    }
                    idx[j][k] = decode_flat(tx, ty, W, H, npx_src);
                }
            }
            if (KIND == 1 && MODE != 2 && redo_bits) qpos = atomicAdd(qn, 1);  // consumed in the next iteration
            base_idx = base;
        } else if (base_px < 0) {
            break;  // nothing left in any stage
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// geo_fast_body — the pixel loop on DOUBLED coordinates (affine always; projective when the denominator keeps one
// sign and a moderate range over the frame window, see geo_fast_mode).
//
// With T = 2*v + 1 (+ delta) + 1.5*2^20 in the 2^-32 fixed-point layout of jsnum.cuh, n = hi(T) - HG_HI_ZERO is
// floor(2v + 1), so
//     Math.round(v) = floor(v + 1/2) = n >> 1            floor(v) = (n - 1) >> 1
//     0 <= v < W  (H.js:1001, on the unrounded v)   <=>   1 <= n <= 2W   <=>   (unsigned)(n - 1) < 2W
// and every decision of the loop (bounds at integers, rounding at half-integers) sits at an INTEGER of 2v: "the
// quotient is within delta of a decision boundary" is "frac(2v + 1 + delta) < 2 delta", one zone instead of two.
//   affine      v is formed exactly as the reference does (jsnum.cuh) and T = fma_rd(v, 2, magic) is one more exact
//               step (the add rounds down onto a grid that holds every integer): no near test at all.
//   projective  the numerators use doubled coefficients (an exact scaling) and ONE fma each with a per-row constant —
//               a different association than the reference's (and one add per pixel along the quad), off by <= 8 ulp of the
//               largest term, which geo_fast_mode bounds below 2^-25 pixel — and the reciprocal of the denominator (MUFU.RCP64H + one
//               Newton step, 2^-39.9 relative).  Total error of 2v < 2^-20.9 < delta = 2^-19.  The eight fractions of a
//               quad are tested with ONE compare of their minimum (a three-input minimum per pixel); only a quad that has a
//               fraction below 2 delta works out which of its pixels are flagged.  Flagged pixels go to the warp queue and
//               are redone by geo_flush_queue with the reference's own arithmetic.
// Two-stage software pipeline, unrolled by two row groups: a group's gathers are issued right after its
// arithmetic and stored one group later, so the loads of 8 pixels per thread stay in flight across a full group of
// arithmetic of the same warp; the bounds predicate guards the load directly (no sentinel round trip).
#define HG_NEAR_DELTA2 8192u  // 2^13 * 2^-32 = 2^-19 in units of 2v

// The magic addends of the doubled-coordinate decode, read from the constant bank: a DFMA takes a constant-bank operand
// directly, a 64-bit literal with a non-zero low word is two moves per row under the kernels' register caps.
__constant__ double hg_mg_affine = HG_MAGIC + 1.0;
__constant__ double hg_mg_projective = HG_MAGIC + 1.0 + (double)HG_NEAR_DELTA2 / 4294967296.0;

struct GeoGroup {
    uint32_t px[GEO_ROWS_PER_THREAD][4];
    int base;        // first row of the group (-1: empty)
    unsigned redo;   // pixels whose quotient must be resolved exactly
    int qpos;        // queue slot reserved for them
};

template <int KIND>
struct GeoFastCtx {
    const uint32_t *src;
    unsigned W, H, W2, H2, npx;
    unsigned Wi, Hi;         // 2W - 3, 2H - 3 (0 for a 1-pixel dimension): the strictly-inside range of the end-pixel test
    unsigned kflat;          // flat = (hy >> 1) * W + (hx >> 1) - kflat
    unsigned nkflat;         // -kflat: (hx >> 1) + nkflat is one LEA.HI
    const uint32_t *srcv;    // src again, pinned to vector registers (the address multiply-add then takes an immediate 4)
    unsigned last[GEO_ROWS_PER_THREAD];  // flat index of the row's first pixel in the previous group (L2 prefetch stride)
    double xs[4];
    double c0, c1, c2, c3, c4, c5, c6, c7;  // affine: m0..m5; projective: 2h0, 2h1, 2h2, 2h3, 2h4, 2h5, h6, h7
    int xOff, yOff, s;
    uint32_t ring;           // ASYNC: shared-memory byte address of this thread's word in slot 0 of the gather ring
};

// ASYNC variant of the pixel loop (warp_inverse_geo_async_kernel): the gathers are 4-byte asynchronous copies global ->
// shared (cp.async) into a ring of two row groups, words [slot][row j][pixel k][thread] (the 32 lanes of one copy write
// 32 consecutive words), committed as one group per row group and awaited only where the row group is stored.  ptxas
// tracks every LDG of the register variant with ONE scoreboard barrier, so storing a row group waits for the gathers of
// the NEXT one as well (issued a moment before); asynchronous copies carry no register scoreboard.  A pixel outside the
// image is a copy of zero source bytes (the slot is zero-filled).
constexpr uint32_t GEO_RING_K = 4u * GEO_THREADS;                          // bytes between pixel k and k+1
constexpr uint32_t GEO_RING_J = 4u * GEO_RING_K;                           // bytes between row j and j+1
constexpr uint32_t GEO_RING_SLOT = GEO_ROWS_PER_THREAD * GEO_RING_J;       // bytes per ring slot

__device__ __forceinline__ void geo_copy4(uint32_t smem_dst, const uint32_t *src, bool take)
{
    const int n = take ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_dst), "l"(src), "r"(n) : "memory");
}

// arithmetic + gathers of one row group
template <int KIND, bool ASYNC = false>
__device__ __forceinline__ void geo_fast_issue(GeoFastCtx<KIND> &C, GeoGroup &g, int base, int *qn, uint32_t slot = 0u)
{
    constexpr int R = GEO_ROWS_PER_THREAD;
    const double MG = (KIND == 0) ? hg_mg_affine : hg_mg_projective;
    g.base = base;
    g.redo = 0u;
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const double y = (double)(C.yOff + base + C.s * j);
        double r0 = 0.0, r1 = 0.0;
        if (KIND == 0) {
            r0 = __dmul_rn(C.c2, y);
            r1 = __dmul_rn(C.c3, y);
        }
        // projective: numerators / denominator of the quad's first pixel — ONE fma each: the thread's x never changes, so
        // 2 h0 x + 2 h2, 2 h3 x + 2 h5 and h6 x + 1 are formed once per thread (C.xs[1..3]) —, then one add per pixel (x
        // advances by 1)
        double nx = 0.0, ny = 0.0, dn = 0.0;
        if (KIND == 1) {
            nx = __fma_rn(C.c1, y, C.xs[1]);
            ny = __fma_rn(C.c4, y, C.xs[2]);
            dn = __fma_rn(C.c7, y, C.xs[3]);
        }
        unsigned hx[4], hy[4];
        unsigned lx[4], ly[4], lomin = 0xFFFFFFFFu;   // projective: fractions of 2v + 1 + delta, and the smallest of the quad's eight
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double tx, ty;
            if (KIND == 0) {
                tx = __fma_rd(affine_coord_exact(C.c0, C.xs[k], r0, C.c4), 2.0, MG);
                ty = __fma_rd(affine_coord_exact(C.c1, C.xs[k], r1, C.c5), 2.0, MG);
            } else {
                const double rc = rcp_newton1(dn);
                tx = __fma_rn(nx, rc, MG);
                ty = __fma_rn(ny, rc, MG);
                if (k < 3) {
                    nx = __dadd_rn(nx, C.c0);
                    ny = __dadd_rn(ny, C.c3);
                    dn = __dadd_rn(dn, C.c6);
                }
                lx[k] = (unsigned)__double2loint(tx);
                ly[k] = (unsigned)__double2loint(ty);
                lomin = __vimin3_u32(lomin, lx[k], ly[k]);   // one three-input minimum per pixel
            }
            hx[k] = (unsigned)__double2hiint(tx);
            hy[k] = (unsigned)__double2hiint(ty);
        }
        // "within delta of a decision boundary" per QUAD first: one compare for eight fractions; the per-pixel flags are
        // worked out only for the (rare) quad that has one
        if (KIND == 1 && lomin < 2u * HG_NEAR_DELTA2) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool again = (lx[k] < 2u * HG_NEAR_DELTA2) | (ly[k] < 2u * HG_NEAR_DELTA2);
                g.redo |= again ? (1u << (4 * j + k)) : 0u;
            }
        }
        // Along an output row both source coordinates are monotone in x (the denominator keeps its sign), so when the
        // quad's two END pixels lie at least one half-pixel step inside the image (2 <= n <= 2W - 2, same for y) the
        // two middle ones lie inside too — also after the <= 1 step an approximate, flagged quotient may be off —
        // and their flat index is a real pixel: no per-pixel test, no predicate on the load.
#ifndef HG_GEO_NO_PREFETCH
        {
            // L2 prefetch one group ahead: the row's first pixel moves by an almost constant flat stride from group to
            // group, so "this group's index + the stride since the last group" names the cache line the next group
            // will gather from.  A hint only: a wrong guess costs one useless line, never a wrong pixel.
            const unsigned f0 = (hy[0] >> 1) * C.W + ((hx[0] >> 1) + C.nkflat);
            const unsigned pf = f0 + (f0 - C.last[j]);
            C.last[j] = f0;
            // only when the quad reads along one source row (a rotated map walks down a column: one line per pixel,
            // where a single hint per quad is noise)
            if ((pf < C.npx) & ((hy[0] ^ hy[3]) < 2u)) asm volatile("prefetch.global.L2 [%0];" ::"l"(C.src + pf));
        }
#endif
        const unsigned cz = (unsigned)(HG_HI_ZERO + 2);
        const bool ends_inside = ((hx[0] - cz) < C.Wi) & ((hy[0] - cz) < C.Hi) & ((hx[3] - cz) < C.Wi) & ((hy[3] - cz) < C.Hi);
        // decided per warp (the lanes that run the body stay together), so it is a real branch, not predication of
        // both variants
        if (__all_sync(__activemask(), ends_inside)) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t *p = C.srcv + ((hy[k] >> 1) * C.W + ((hx[k] >> 1) + C.nkflat));
                if (ASYNC) geo_copy4(C.ring + slot + GEO_RING_J * j + GEO_RING_K * k, p, true);
                else g.px[j][k] = __ldg(p);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned flat = (hy[k] >> 1) * C.W + ((hx[k] >> 1) + C.nkflat);
                // 0 <= v < W on the unrounded coordinate, and the flat index inside the image (Q2: past the end reads 0)
                const bool in = ((hx[k] - (unsigned)(HG_HI_ZERO + 1)) < C.W2) & ((hy[k] - (unsigned)(HG_HI_ZERO + 1)) < C.H2) &
                                (flat < C.npx);
                if (ASYNC) {
                    geo_copy4(C.ring + slot + GEO_RING_J * j + GEO_RING_K * k, C.src + (in ? flat : 0u), in);
                } else {
                    uint32_t v = 0u;
                    if (in) v = __ldg(C.src + flat);
                    g.px[j][k] = v;
                }
            }
        }
    }
    if (ASYNC) asm volatile("cp.async.commit_group;" ::: "memory");
    g.qpos = 0;
    if (KIND == 1 && g.redo) g.qpos = atomicAdd(qn, 1);  // consumed when the group retires
}

// queue entry (or in-place exact resolution when the queue is full), then the stores of one row group
template <int KIND, bool ASYNC = false>
__device__ __forceinline__ void geo_fast_retire(const GeoFrame &F, const double (&m)[8], const GeoFastCtx<KIND> &C, GeoGroup &g,
                                                int x_first, unsigned mask, uint2 *q, const uint32_t *slot = nullptr, int pending = 0)
{
    constexpr int R = GEO_ROWS_PER_THREAD;
    if (g.base < 0) return;
    if (ASYNC) {
        // the copies of this row group have landed once at most `pending` newer groups are still in flight
        if (pending) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
        for (int j = 0; j < R; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) g.px[j][k] = slot[(j * 4 + k) * GEO_THREADS];
    }
    if (KIND == 1 && g.redo) {
        if (g.qpos < GEO_QCAP) {
            q[g.qpos] = make_uint2((unsigned)(x_first + 4), (g.redo << 17) | (unsigned)g.base);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                unsigned rows = (g.redo >> k) & 0x11111111u;
                while (rows) {
                    const int j = (__ffs((int)rows) - 1) >> 2;
                    rows &= rows - 1;
                    const double x = (double)(C.xOff + x_first + k), y = (double)(C.yOff + g.base + C.s * j);
                    const double nx = __dadd_rn(__dadd_rn(__dmul_rn(m[0], x), __dmul_rn(m[1], y)), m[2]);
                    const double ny = __dadd_rn(__dadd_rn(__dmul_rn(m[3], x), __dmul_rn(m[4], y)), m[5]);
                    const double dn = __dadd_rn(__dadd_rn(__dmul_rn(m[6], x), __dmul_rn(m[7], y)), 1.0);
                    const uint32_t v = ldg_or_zero(C.src, decode_flat(exact_quotient_magic(nx, dn), exact_quotient_magic(ny, dn),
                                                                      C.W, C.H, C.npx));
#pragma unroll
                    for (int jj = 0; jj < R; ++jj)
                        if (jj == j) g.px[jj][k] = v;  // select, no dynamic register indexing
                }
            }
        }
    }
    // pixel offsets in 32 bits: oW * oH < 2^31 (check_window), so the offset of every row that is stored fits an int
    int off = (int)((unsigned)g.base * (unsigned)F.oW) + x_first;
    const int row_step = (int)((unsigned)C.s * (unsigned)F.oW);
#pragma unroll
    for (int j = 0; j < R; ++j) {
        if (g.base + C.s * j < F.oH) {
            uint32_t *dst = F.out + off;
            if (mask == 0xFu) {
                *reinterpret_cast<uint4 *>(dst) = make_uint4(g.px[j][0], g.px[j][1], g.px[j][2], g.px[j][3]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (mask & (1u << k)) dst[k] = g.px[j][k];
            }
        }
        off = (int)((unsigned)off + (unsigned)row_step);
    }
    g.base = -1;
}

template <int KIND, bool ASYNC = false>
__device__ __forceinline__ void geo_fast_body(const GeoFrame &F, const double (&m)[8], int base0, int niter, int s,
                                              int x_first, unsigned mask, uint2 *q, int *qn, int group_rows,
                                              uint32_t *ring = nullptr)
{
    GeoFastCtx<KIND> C;
    C.ring = ASYNC ? (uint32_t)__cvta_generic_to_shared(ring + threadIdx.x) : 0u;
    const uint32_t *ring_a = ring + threadIdx.x, *ring_b = ring + threadIdx.x + GEO_RING_SLOT / 4;
    C.src = F.src;
    C.W = (unsigned)F.W;
    C.H = (unsigned)F.H;
    C.W2 = 2u * C.W;
    C.H2 = 2u * C.H;
    C.Wi = C.W2 >= 3u ? C.W2 - 3u : 0u;
    C.Hi = C.H2 >= 3u ? C.H2 - 3u : 0u;
    // pinned to registers: under the kernels' register caps ptxas otherwise re-derives both (subtract, compare, select) in
    // every row of the loop
    asm volatile("" : "+r"(C.Wi), "+r"(C.Hi));
    C.npx = C.W * C.H;
    C.kflat = (unsigned)(HG_HI_ZERO >> 1) * (C.W + 1u);
    C.nkflat = 0u - C.kflat;
    asm volatile("" : "+r"(C.nkflat));   // one live constant: ptxas otherwise re-derives it from W in every row
    C.srcv = F.src;
    asm volatile("" : "+l"(C.srcv));
    C.xOff = F.xOff;
    C.yOff = F.yOff;
    C.s = s;
#pragma unroll
    for (int j = 0; j < GEO_ROWS_PER_THREAD; ++j) C.last[j] = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < 4; ++k) C.xs[k] = (double)(F.xOff + x_first + (KIND == 0 ? k : 0));  // projective: replaced below
    if (KIND == 0) {
        C.c0 = m[0]; C.c1 = m[1]; C.c2 = m[2]; C.c3 = m[3]; C.c4 = m[4]; C.c5 = m[5]; C.c6 = 0.0; C.c7 = 0.0;
    } else {
        C.c0 = 2.0 * m[0]; C.c1 = 2.0 * m[1]; C.c2 = 2.0 * m[2]; C.c3 = 2.0 * m[3]; C.c4 = 2.0 * m[4]; C.c5 = 2.0 * m[5];
        C.c6 = m[6]; C.c7 = m[7];
        // keep the doubled coefficients in registers (the compiler would otherwise re-derive them from the kernel
        // parameters inside the loop)
        asm volatile("" : "+d"(C.c0), "+d"(C.c1), "+d"(C.c2), "+d"(C.c3), "+d"(C.c4), "+d"(C.c5));
        // the x part of the numerators and of the denominator, once per thread (two roundings per value like the per-row
        // association it replaces: the error bound of geo_fast_mode is unchanged)
        const double x0 = C.xs[0];
        C.xs[1] = __fma_rn(C.c0, x0, C.c2);
        C.xs[2] = __fma_rn(C.c3, x0, C.c5);
        C.xs[3] = __fma_rn(C.c6, x0, 1.0);
    }
    GeoGroup ga, gb;
    ga.base = gb.base = -1;
    const int oH = F.oH;
#pragma unroll 1
    for (int it = 0; it < niter; it += 2) {
        const int b0 = base0 + it * group_rows, b1 = b0 + group_rows;
        if (b0 >= oH) break;
        geo_fast_issue<KIND, ASYNC>(C, ga, b0, qn, 0u);
        geo_fast_retire<KIND, ASYNC>(F, m, C, gb, x_first, mask, q, ring_b, 1);   // ga's copies stay in flight
        if (it + 1 < niter && b1 < oH) {
            geo_fast_issue<KIND, ASYNC>(C, gb, b1, qn, GEO_RING_SLOT);
            geo_fast_retire<KIND, ASYNC>(F, m, C, ga, x_first, mask, q, ring_a, 1);
        }
    }
    geo_fast_retire<KIND, ASYNC>(F, m, C, ga, x_first, mask, q, ring_a, 0);
    geo_fast_retire<KIND, ASYNC>(F, m, C, gb, x_first, mask, q, ring_b, 0);
}

// CTA-uniform: may the projective frame run geo_fast_body?  The denominator h6 x + h7 y + 1 is affine in (x, y), so
// its range over the frame window is spanned by the four corners.
__device__ __forceinline__ bool geo_fast_mode(const GeoFrame &F, const double (&m)[8])
{
    if (m[6] == 0.0 && m[7] == 0.0) return false;  // denominator == 1: geo_tile_body<1,2> is exact without a reciprocal
    const double X0 = (double)(F.xOff - 3), X1 = (double)(F.xOff + F.oW + 3);
    const double Y0 = (double)F.yOff, Y1 = (double)(F.yOff + F.oH + 15);
    double dmin = 1e300, dmax = 0.0;
    bool pos = true, neg = true;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const double dn = m[6] * ((c & 1) ? X1 : X0) + m[7] * ((c & 2) ? Y1 : Y0) + 1.0;
        pos = pos && (dn > 0.0);
        neg = neg && (dn < 0.0);
        dmin = fmin(dmin, fabs(dn));
        dmax = fmax(dmax, fabs(dn));
    }
    if (!(pos || neg) || !(dmin >= 0.015625) || !(dmax <= 64.0)) return false;  // also false for NaN
    const double Xm = fmax(fabs(X0), fabs(X1)), Ym = fmax(fabs(Y0), fabs(Y1)), big = 16777216.0 * dmin;  // 2^24 dmin
    return (fabs(m[0]) * Xm + fabs(m[1]) * Ym + fabs(m[2]) < big) && (fabs(m[3]) * Xm + fabs(m[4]) * Ym + fabs(m[5]) < big) &&
           (fabs(m[6]) * Xm + fabs(m[7]) * Ym + 1.0 < 256.0 * dmin);
}

// Is RN(N / D) >= b, decided exactly WITHOUT the division?  b is a non-zero multiple of 1/2 with |b| < 2^19 (an even
// mantissa), D is finite, non-zero and of moderate size, N finite.
//   RN(Q) >= b  <=>  Q >= b - h,   h = half the gap between b and the double just below it (at the exact midpoint the
//   tie goes to the even neighbour, which is b);  with the sign of D:  N - b D + h D >= 0  (D > 0),  <= 0  (D < 0).
//   r = fma(-b, D, N) is exact whenever it is small enough to matter: b has <= 20 significant bits, so b*D is a
//   multiple of 2^-72 |N| and a difference below 2^-30 |N| has fewer than 53 bits; a larger |r| dwarfs h|D| <= 2^-52 |bD|
//   and only its sign counts.  t = fma(h, D, r) then carries the exact sign (an fma rounds the exact sum once and
//   never across zero).
__device__ __forceinline__ bool quotient_at_least(double N, double D, double b)
{
    const long long bits = __double_as_longlong(b);
    const int E = (int)((bits >> 52) & 0x7FF);
    const bool pow2_pos = (bits > 0) && ((bits & 0xFFFFFFFFFFFFFLL) == 0);  // gap below a positive power of two is halved
    const double h = __longlong_as_double((long long)(E - 53 - (pow2_pos ? 1 : 0)) << 52);
    const double t = __fma_rn(h, D, __fma_rn(-b, D, N));
    return D > 0.0 ? (t >= 0.0) : (t <= 0.0);
}

// n = floor(2 * RN(N / D) + 1) — the doubled-coordinate integer of geo_fast_body — exactly.  T is the approximate
// 2q + 1 + delta + magic (error < delta): away from an integer of 2q + 1 its integer part is already right; within
// delta of one, the only open question is on which side of b = (n_candidate - 1) / 2 the reference's quotient falls.
__device__ __forceinline__ int exact_doubled_floor(double T, double N, double D)
{
    int n = __double2hiint(T) - HG_HI_ZERO;
    if ((unsigned)__double2loint(T) < 2u * HG_NEAR_DELTA2) {
        const double b = (double)(n - 1) * 0.5;
        bool ge;
        if (b == 0.0) {
            // RN(Q) >= 0 (also true for -0): the sign of Q unless it underflows — then let the division decide
            if (N == 0.0) ge = true;
            else if (fabs(N) < 1e-290) ge = __ddiv_rn(N, D) >= 0.0;
            else ge = (N > 0.0) == (D > 0.0);
        } else {
            ge = quotient_at_least(N, D, b);
        }
        n -= ge ? 0 : 1;
    }
    return n;
}

// Resolve the queued pixels of one warp with the reference's own arithmetic (H.js:1401-1404 followed by the bounds
// test, Math.round and the flat read of H.js:1001-1007), one queue entry per lane, and overwrite them in the output.
// FAST (frames that run geo_fast_body: denominator of one sign and moderate size): the reference's numerators and
// denominator are formed exactly, and instead of two IEEE divisions the side of the one nearby decision boundary is
// settled by quotient_at_least (two fmas).  Otherwise: the divisions themselves.
template <bool FAST>
__device__ __forceinline__ void geo_flush_queue(const GeoFrame &F, const double (&m)[8], const uint2 *q, int n, int s,
                                                int lane)
{
    const unsigned W = (unsigned)F.W, H = (unsigned)F.H;
    const unsigned npx_src = W * H;
    for (int e = lane; e < n; e += 32) {
        const uint2 ent = q[e];
        const int x_first = (int)ent.x - 4, base = (int)(ent.y & 0x1FFFFu);
        unsigned bits = ent.y >> 17;
        while (bits) {
            const int b = __ffs((int)bits) - 1;
            bits &= bits - 1;
            const int xx = x_first + (b & 3), yy = base + s * (b >> 2);
            if (xx < 0 || xx >= F.oW || yy >= F.oH) continue;  // not a pixel of the image
            const double x = (double)(F.xOff + xx), y = (double)(F.yOff + yy);
            const double nx = __dadd_rn(__dadd_rn(__dmul_rn(m[0], x), __dmul_rn(m[1], y)), m[2]);
            const double ny = __dadd_rn(__dadd_rn(__dmul_rn(m[3], x), __dmul_rn(m[4], y)), m[5]);
            const double dn = __dadd_rn(__dadd_rn(__dmul_rn(m[6], x), __dmul_rn(m[7], y)), 1.0);
            uint32_t v;
            if (FAST) {
                const double MG = HG_MAGIC + 1.0 + (double)HG_NEAR_DELTA2 / 4294967296.0;
                const double rc = rcp_newton1(dn);
                const double Tx = __fma_rn(__dadd_rn(nx, nx), rc, MG), Ty = __fma_rn(__dadd_rn(ny, ny), rc, MG);
                const int jx = exact_doubled_floor(Tx, nx, dn), jy = exact_doubled_floor(Ty, ny, dn);
                // Most queued pixels sit EXACTLY on a decision boundary ("nice" points: whole columns of them) and resolve
                // to the value the loop already stored: the exact integers equal the ones an approximate quotient decodes
                // to.  The loop's quotient and this one differ by < 2^-20, so unless one of them lies just BELOW an integer
                // (low word within 4 delta of 2^32: the two could then sit on different sides of it) both decode to the
                // same integer part, and the provisional pixel is the reference's — no gather, no store.
                const unsigned top = 0u - 4u * HG_NEAR_DELTA2;
                if (jx == __double2hiint(Tx) - HG_HI_ZERO && jy == __double2hiint(Ty) - HG_HI_ZERO &&
                    (unsigned)__double2loint(Tx) < top && (unsigned)__double2loint(Ty) < top)
                    continue;
                const unsigned flat = (unsigned)(jy >> 1) * W + (unsigned)(jx >> 1);
                const bool in = ((unsigned)(jx - 1) < 2u * W) & ((unsigned)(jy - 1) < 2u * H) & (flat < npx_src);
                v = in ? __ldg(F.src + flat) : 0u;
            } else {
                v = ldg_or_zero(F.src, decode_flat(exact_quotient_magic(nx, dn), exact_quotient_magic(ny, dn), W, H, npx_src));
            }
            F.out[(long long)yy * F.oW + xx] = v;
        }
    }
}

// The default kernel: one CTA per (tile, frame), source pixels gathered directly from global memory through L1 / L2
// (geo_fast_body; geo_tile_body for the projective frames geo_fast_mode rejects).
template <int KIND, bool ASYNC>
__device__ __forceinline__ void warp_inverse_geo_impl(const GeoParams &P, uint32_t *ring);

template <int KIND>
__global__ void __launch_bounds__(GEO_THREADS, KIND == 1 ? HG_GEO_MINB_PROJ : HG_GEO_MINB) warp_inverse_geo_kernel(const GeoParams P)
{
    warp_inverse_geo_impl<KIND, false>(P, nullptr);
}

// the affine kernel compiled for the tall thread layout (see HG_GEO_MINB_TALL)
__global__ void __launch_bounds__(GEO_THREADS, HG_GEO_MINB_TALL) warp_inverse_geo_affine_tall_kernel(const GeoParams P)
{
    warp_inverse_geo_impl<0, false>(P, nullptr);
}

// the same kernel with asynchronous gathers (see GeoFastCtx): opt-in / default per launch_geo
template <int KIND>
__global__ void __launch_bounds__(GEO_THREADS, KIND == 1 ? HG_GEO_MINB_PROJ : HG_GEO_MINB) warp_inverse_geo_async_kernel(const GeoParams P)
{
    __shared__ uint32_t s_ring[2 * GEO_RING_SLOT / 4];
    warp_inverse_geo_impl<KIND, true>(P, s_ring);
}

template <int KIND, bool ASYNC>
__device__ __forceinline__ void warp_inverse_geo_impl(const GeoParams &P, uint32_t *ring)
{
    // per-warp queue of pixels whose quotient must be resolved exactly (projective only)
    __shared__ uint2 s_q[KIND == 1 ? GEO_THREADS / 32 : 1][KIND == 1 ? GEO_QCAP : 1];
    __shared__ int s_qn[GEO_THREADS / 32];
    const int warp_id = threadIdx.x >> 5, lane_id = threadIdx.x & 31;
    if (lane_id == 0) s_qn[warp_id] = 0;
    __syncwarp();

    const GeoFrame F = P.many ? P.many[blockIdx.y] : P.one;
    double m[8];
    if (P.mats_dev) {
        if (KIND == 0) {
            const float *mf = (const float *)P.mats_dev + 6 * (size_t)blockIdx.y;
#pragma unroll
            for (int k = 0; k < 6; ++k) m[k] = (double)__ldg(mf + k);
            m[6] = m[7] = 0.0;
        } else {
            const double *md = (const double *)P.mats_dev + 8 * (size_t)blockIdx.y;
#pragma unroll
            for (int k = 0; k < 8; ++k) m[k] = __ldg(md + k);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = P.mat_val[k];
    }

    const int oW = F.oW, oH = F.oH;
    const int ltx = P.ltx, tile_quads = 1 << ltx, group_rows = geo_group_rows(ltx);
    const int tiles_x = geo_tiles_x(oW, tile_quads);
    const int tile_y = blockIdx.x / tiles_x;
    const int tile_x = blockIdx.x - tile_y * tiles_x;
    const int row0 = tile_y * (group_rows * P.niter);
    if (row0 >= oH) return;  // grid is sized for the largest frame of the batch

    // rows with equal flat alignment repeat with period s = 4 / gcd(oW mod 4, 4)
    const int sl = (oW & 3) == 0 ? 0 : ((oW & 1) ? 2 : 1);  // log2(s)
    const int s = 1 << sl;
    const int tx = threadIdx.x & (tile_quads - 1), ty = threadIdx.x >> ltx;
    const int base = row0 + (ty >> sl) * (s * GEO_ROWS_PER_THREAD) + (ty & (s - 1));
    const int shift = (int)(((unsigned)base * (unsigned)oW) & 3u);
    const int x_first = 4 * (tile_x * tile_quads + tx) - shift;
    const bool active = (base < oH) && (x_first < oW);
    unsigned mask = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (x_first + k >= 0 && x_first + k < oW) mask |= 1u << k;
    uint2 *q = s_q[KIND == 1 ? warp_id : 0];
    int *qn = &s_qn[warp_id];

    const bool fast = (KIND == 1) && geo_fast_mode(F, m);  // CTA-uniform: from the matrix and the frame window
    if (active) {
        if (KIND == 0) {
            geo_fast_body<0, ASYNC>(F, m, base, P.niter, s, x_first, mask, q, qn, group_rows, ring);
        } else {
            if (m[6] == 0.0 && m[7] == 0.0) geo_tile_body<1, 2>(F, m, base, P.niter, s, x_first, mask, q, qn, group_rows);
            else if (fast) geo_fast_body<1, ASYNC>(F, m, base, P.niter, s, x_first, mask, q, qn, group_rows, ring);
            else geo_tile_body<1, 0>(F, m, base, P.niter, s, x_first, mask, q, qn, group_rows);
        }
    }
    if (KIND == 1) {
        // every lane of the warp gets here; the barrier also orders the provisional stores above before the
        // corrected ones below
        __syncwarp();
        const int n = min(*qn, GEO_QCAP);
        if (n > 0) {
            if (fast) geo_flush_queue<true>(F, m, q, n, s, lane_id);
            else geo_flush_queue<false>(F, m, q, n, s, lane_id);
        }
    }
}
}  // namespace hg

#include "warp_geo_staged.cuh"
