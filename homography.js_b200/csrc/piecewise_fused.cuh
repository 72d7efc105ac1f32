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
// (~1/64 of the map's size; a few entries per bin for any non-folded mesh).  The overlaps are then resolved — highest id
// wins — and each bin is run-length encoded as a 64-bit start mask + up to 8 int16 ids.  Two ways to get there: meshes of
// up to 2,048 triangles take ONE pass per band of map rows that keeps the band's bins in shared memory and writes the run
// records directly (pw_band_bins_kernel); finer meshes bin into global memory (pw_span_bin_kernel) and encode in a second
// pass (pw_bin_runs_kernel, one thread per bin).  The warp kernel finds t with two popcounts, then does what
// H.js:1046-1052 does: inverse 2x3 of the triangle, window test, Math.round, flat gather, 128-bit store.
//
// Exactness: every quirk of the reference's map (no x offset -> spans spilling into the next row, negative
// relative fill indices landing at the END of the map, last-writer-wins overlaps, int16 wrap of ids) is inherited
// from the exact interval computation.  A frame that cannot be represented (a bin with more than PW_BIN_CAP
// entries or runs, an interval crossing more than PW_MAX_PIECES rows or leaving the rows the band pass assumed it could reach,
// >= 2^17 triangles) raises a status flag and is redone by the general map-based path — never approximated.
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
    const float *inv;      // n_tris * 8 floats: the f32-rounded inverse matrices, one 32-byte record each
    const int2 *yr;        // n_tris row ranges [y0, y_end) of the triangles (pw_band_bins_kernel scans these, not the records)
    unsigned *bin_cnt;     // oH * bins_x x {entry counter, highest whole-bin id + 1} (zeroed before the span kernel)
    unsigned *bin_ent;     // oH * bins_x * PW_BIN_CAP packed entries  (t << 14 | c1 << 7 | c0)
    uint4 *bin_run;        // oH * bins_x run records (pw_bin_runs_kernel): what the pixel kernel reads
    int *status;           // bit 0: not representable -> redo with the general path
    int W, H, xOff, yOff, oW, oH, minSrcX, minSrcY, n_tris, bins_x;
};

// `lpt` lanes per triangle (a power of two: 8 .. 512), chosen on the host from the mesh so that a triangle's rows fill
// its lanes about three times over — short triangles of fine meshes share a warp, tall triangles of coarse meshes spread
// over several warps.  Lane l of a triangle takes its rows l, l + lpt, ...
// Per row: the exact interval [S, E) of TypedArray.fill (IEEE edge intersections, Math.round, relative-index clamping:
// H.js:1111-1126, 1172-1197), cut at map-row boundaries, one entry per 64-column bin it touches.  All flat indices are
// below 2^31 (checked on the host), so the piece arithmetic is 32-bit.
// Bin state: two words per bin — [2 bin] a counter of entries, [2 bin + 1] the highest id + 1 among the intervals that cover
// the WHOLE bin (0 = none).  A wide interval (coarse meshes: seven bins per row and triangle) covers its interior bins
// completely: those need no entry, only a fire-and-forget atomic max (RED) — no slot to wait for.  Entries (slot
// reservation with a returning atomic, then a store that depends on it) are left for the two ends of an interval, and
// their stores are deferred until the next reservation goes out, so the atomic's round trip to L2 runs under the next
// row's edge intersections instead of stalling the lane.
#ifndef HG_SPAN_MINB
#define HG_SPAN_MINB 10
#endif
struct PwPending {
    unsigned *addr;   // entry slot base of the bin (bin * PW_BIN_CAP), nullptr = nothing pending
    unsigned slot, value;
};
__device__ __forceinline__ void pw_pending_flush(PwPending &p, int *status)
{
    if (p.addr) {
        if (p.slot < PW_BIN_CAP) p.addr[p.slot] = p.value;
        else atomicOr(status, 1);
        p.addr = nullptr;
    }
}
template <bool DEFER>
__device__ __forceinline__ void pw_emit_entry(PwPending &p, unsigned *cnt, unsigned *ent, unsigned b, unsigned value, int *status)
{
    if (!DEFER) {
        const unsigned slot = atomicAdd(cnt + 2u * b, 1u);
        if (slot < PW_BIN_CAP) ent[(size_t)b * PW_BIN_CAP + slot] = value;
        else atomicOr(status, 1);
        return;
    }
    pw_pending_flush(p, status);
    p.slot = atomicAdd(cnt + 2u * b, 1u);
    p.addr = ent + (size_t)b * PW_BIN_CAP;
    p.value = value;
}

// DEFER: entry stores wait until the next reservation goes out (coarse meshes: few, tall triangles, the kernel is bound by
// the latency of its atomics — 57 -> 37 us per 16 4K frames of the 162-triangle mesh); fine meshes have thousands of
// triangles in flight to cover that latency and measured 15 % faster storing at once.
template <bool DEFER>
__global__ void __launch_bounds__(128, HG_SPAN_MINB) pw_span_bin_kernel(const FusedFrame *frames, int lpt_log2)
{
    const FusedFrame F = frames[blockIdx.y];   // by value: the atomics below may alias the descriptor as far as the compiler knows
    const unsigned gid = blockIdx.x * 128u + threadIdx.x;
    const int t = (int)(gid >> lpt_log2);
    if (t >= F.n_tris) return;
    const int lane = (int)(gid & ((1u << lpt_log2) - 1u)), stride = 1 << lpt_log2;
    TriRec r;
    {
        // the record through the read-only path into locals: the entry stores below may alias it as far as the compiler
        // knows, and every row would otherwise wait for twelve edge doubles re-read behind them (40 warps per SM are worth
        // more than keeping all of them in registers: the launch bound lets the compiler decide which to re-fetch)
        const TriRec *__restrict__ g = F.rec + t;
#pragma unroll
        for (int e = 0; e < 3; ++e) { r.m[e] = __ldg(&g->m[e]); r.b[e] = __ldg(&g->b[e]); r.lo[e] = __ldg(&g->lo[e]); r.hi[e] = __ldg(&g->hi[e]); }
        r.maxY = __ldg(&g->maxY);
        r.y0 = __ldg(&g->y0);
    }
    const unsigned oW = (unsigned)F.oW, len = oW * (unsigned)F.oH;
    const double mw = (double)F.oW, yoff = (double)F.yOff, dlen = (double)len;
    const double y00 = (double)r.y0, maxY = r.maxY;
    const unsigned tt = (unsigned)t << 14;
    PwPending pa{nullptr, 0u, 0u}, pb{nullptr, 0u, 0u};
    for (int i = lane;; i += stride) {
        const double y = y00 + (double)i;
        if (!(y < maxY)) break;  // also ends on NaN
        double xo, xd;
        predict_x_limits(r, y, xo, xd);
        const double rowbase = __dmul_rn(__dsub_rn(y, yoff), mw);
        const double rel0 = __dadd_rn(rowbase, js_round(xo)), rel1 = __dadd_rn(rowbase, js_round(xd));
        unsigned k0, k1;
        if (rel0 >= 0.0 && rel0 < dlen && rel1 >= 0.0) {
            // the usual case: both relative indices are non-negative integers (a product and a sum of small integers),
            // fill() clamps the end to the length
            k0 = (unsigned)(int)rel0;
            k1 = rel1 < dlen ? (unsigned)(int)rel1 : len;
        } else {  // negative (counted from the end), infinite or NaN indices: the general rule
            k0 = (unsigned)js_fill_bound(rel0, (long long)len);
            k1 = (unsigned)js_fill_bound(rel1, (long long)len);
        }
        if (k0 >= k1) continue;
        unsigned row = k0 / oW, c0 = k0 - row * oW;
        int pieces = 0;
        while (k0 < k1) {
            if (++pieces > PW_MAX_PIECES) { atomicOr(F.status, 1); break; }
            const unsigned room = oW - c0, want = k1 - k0, n = want < room ? want : room;
            const unsigned c1 = c0 + n;  // the piece covers columns [c0, c1) of map row `row`
            const unsigned b_first = c0 / PW_BIN_W, b_last = (c1 - 1u) / PW_BIN_W;
            unsigned *cnt = F.bin_cnt + 2 * (size_t)row * F.bins_x;
            unsigned *ent = F.bin_ent + (size_t)row * F.bins_x * PW_BIN_CAP;
            const unsigned lo = c0 - b_first * PW_BIN_W, hi = c1 - b_last * PW_BIN_W;
            if (DEFER) {
                if (b_first == b_last) {
                    if (lo == 0u && hi == (unsigned)PW_BIN_W) atomicMax(cnt + 2u * b_first + 1u, (unsigned)t + 1u);
                    else pw_emit_entry<DEFER>(pa, cnt, ent, b_first, tt | (hi << 7) | lo, F.status);
                } else {
                    // the first and the last bin of the piece are partial as a rule (an entry each), the bins between them are
                    // covered completely
                    if (lo == 0u) atomicMax(cnt + 2u * b_first + 1u, (unsigned)t + 1u);
                    else pw_emit_entry<DEFER>(pa, cnt, ent, b_first, tt | ((unsigned)PW_BIN_W << 7) | lo, F.status);
                    for (unsigned b = b_first + 1u; b < b_last; ++b) atomicMax(cnt + 2u * b + 1u, (unsigned)t + 1u);
                    if (hi == (unsigned)PW_BIN_W) atomicMax(cnt + 2u * b_last + 1u, (unsigned)t + 1u);
                    else pw_emit_entry<DEFER>(pb, cnt, ent, b_last, tt | (hi << 7), F.status);
                }
            } else {
                // The (at most two) entries of the piece are noted first and their slots reserved at ONE place, both
                // reservations in front of both stores: the lanes of a warp take different branches above (one bin, partial
                // first bin, partial last bin), and a reservation inside each branch runs with the few lanes of that branch
                // and a round trip to L2 of its own — up to three per row instead of one.
                unsigned e_bin[2] = {0u, 0u}, e_val[2] = {0u, 0u};
                int ne = 0;
                if (b_first == b_last) {
                    if (lo == 0u && hi == (unsigned)PW_BIN_W) atomicMax(cnt + 2u * b_first + 1u, (unsigned)t + 1u);
                    else { e_bin[0] = b_first; e_val[0] = tt | (hi << 7) | lo; ne = 1; }
                } else {
                    if (lo == 0u) atomicMax(cnt + 2u * b_first + 1u, (unsigned)t + 1u);
                    else { e_bin[0] = b_first; e_val[0] = tt | ((unsigned)PW_BIN_W << 7) | lo; ne = 1; }
                    for (unsigned b = b_first + 1u; b < b_last; ++b) atomicMax(cnt + 2u * b + 1u, (unsigned)t + 1u);
                    if (hi == (unsigned)PW_BIN_W) atomicMax(cnt + 2u * b_last + 1u, (unsigned)t + 1u);
                    else {
                        if (ne == 0) { e_bin[0] = b_last; e_val[0] = tt | (hi << 7); }
                        else { e_bin[1] = b_last; e_val[1] = tt | (hi << 7); }
                        ++ne;
                    }
                }
                unsigned s0 = 0u, s1 = 0u;
                if (ne > 0) s0 = atomicAdd(cnt + 2u * e_bin[0], 1u);
                if (ne > 1) s1 = atomicAdd(cnt + 2u * e_bin[1], 1u);
                if (ne > 0) {
                    if (s0 < (unsigned)PW_BIN_CAP) ent[(size_t)e_bin[0] * PW_BIN_CAP + s0] = e_val[0];
                    else atomicOr(F.status, 1);
                }
                if (ne > 1) {
                    if (s1 < (unsigned)PW_BIN_CAP) ent[(size_t)e_bin[1] * PW_BIN_CAP + s1] = e_val[1];
                    else atomicOr(F.status, 1);
                }
            }
            k0 += n;
            ++row;
            c0 = 0u;
        }
    }
    pw_pending_flush(pa, F.status);
    pw_pending_flush(pb, F.status);
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
// flags the frame for the general path.  Bins with no entry or a single entry (most bins of a coarse mesh) take a
// short cut.  (Measured and dropped: sorting the entries by id through a register network and painting them from the
// highest id down with 64-bit column masks — 33 % slower than scanning the entries per candidate column: bins hold two
// or three entries, the fixed cost of eight-wide masks does not pay.)
constexpr int PW_RUN_CAP = 8;

// the run record of one bin from its span entries (`ent`: 16-byte aligned, global or shared memory), their count and the
// highest id among the intervals that cover the whole bin (-1: none)
__device__ __forceinline__ void pwf_make_record_from(const unsigned *ent_p, unsigned cnt_raw, int full_raw, int n_tris, int *status,
                                                     uint4 &rec0, uint4 &rec1)
{
    const unsigned cnt = min(cnt_raw, (unsigned)PW_BIN_CAP);
    unsigned long long mask = 1ull;
    unsigned ids[PW_RUN_CAP / 2] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};  // int16 pairs, -1 = no triangle
    if (cnt == 0u) {
        ids[0] = 0xFFFF0000u | ((unsigned)pwf_map_id(full_raw, n_tris) & 0xFFFFu);
    } else if (cnt == 1u && full_raw < 0) {
        const unsigned e = ent_p[0];
        const unsigned lo = e & 127u, hi = (e >> 7) & 127u;
        const unsigned id = (unsigned)pwf_map_id((int)(e >> 14), n_tris) & 0xFFFFu;
        if (id != 0xFFFFu) {
            // runs: [nothing, 0..lo) [t, lo..hi) [nothing, hi..64)
            if (lo == 0u) {
                ids[0] = 0xFFFF0000u | id;
            } else {
                mask |= 1ull << lo;
                ids[0] = (id << 16) | 0xFFFFu;
            }
            if (hi < 64u) mask |= 1ull << hi;  // its id (-1) is already in place
        }
    } else {
        unsigned ent[PW_BIN_CAP];
        {
            const uint4 *pe = reinterpret_cast<const uint4 *>(ent_p);
            const uint4 a = pe[0], b = cnt > 4 ? pe[1] : make_uint4(0, 0, 0, 0);
            ent[0] = a.x; ent[1] = a.y; ent[2] = a.z; ent[3] = a.w; ent[4] = b.x; ent[5] = b.y; ent[6] = b.z; ent[7] = b.w;
        }
        // candidate run starts: column 0 and every interval end point inside the bin.  Typical bins hold 2-3 entries: the
        // loops stop at cnt instead of running predicated over all PW_BIN_CAP slots
        unsigned long long cand = 1ull;
#pragma unroll
        for (int e = 0; e < PW_BIN_CAP; ++e) {
            if ((unsigned)e >= cnt) break;
            const unsigned lo = ent[e] & 127u, hi = (ent[e] >> 7) & 127u;
            cand |= 1ull << lo;          // lo <= 63
            if (hi < 64u) cand |= 1ull << hi;
        }
        mask = 0ull;
        int runs = 0, prev = -2;
        while (cand) {
            const int c = __ffsll((long long)cand) - 1;
            cand &= cand - 1;
            int raw = full_raw;   // what covers the whole bin covers this column
#pragma unroll
            for (int e = 0; e < PW_BIN_CAP; ++e) {
                if ((unsigned)e >= cnt) break;
                // lo <= c < hi  <=>  (unsigned)(c - lo) < (unsigned)(hi - lo)
                const int lo = (int)(ent[e] & 127u), hi = (int)((ent[e] >> 7) & 127u);
                if ((unsigned)(c - lo) < (unsigned)(hi - lo)) raw = max(raw, (int)(ent[e] >> 14));
            }
            const int id = pwf_map_id(raw, n_tris);
            if (id != prev) {
                if (runs == PW_RUN_CAP) {
                    atomicOr(status, 1);
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
    }
    rec0 = make_uint4((unsigned)mask, (unsigned)(mask >> 32), 0u, 0u);
    rec1 = make_uint4(ids[0], ids[1], ids[2], ids[3]);
}

__device__ __forceinline__ void pwf_make_record(const FusedFrame &F, size_t bin, uint4 &rec0, uint4 &rec1)
{
    const uint2 st = *reinterpret_cast<const uint2 *>(F.bin_cnt + 2 * bin);   // entries, highest whole-bin id + 1
    pwf_make_record_from(F.bin_ent + bin * PW_BIN_CAP, st.x, (int)st.y - 1, F.n_tris, F.status, rec0, rec1);
}

// the records as a pass of their own: frames whose width is not a multiple of four (their first quads reach back into
// the previous bin, whose record the pixel kernel then needs as well), the general-path comparison runs and the
// first-generation pixel kernel read them from global memory; for the other frames the pixel kernel builds the records
// of its rows itself, straight into shared memory (pw_warp_fused_kernel, records_inline)
__global__ void __launch_bounds__(128) pw_bin_runs_kernel(const FusedFrame *frames)
{
    const FusedFrame F = frames[blockIdx.y];
    const size_t nbins = (size_t)F.bins_x * F.oH;
    const size_t bin = (size_t)blockIdx.x * 128 + threadIdx.x;
    if (bin >= nbins) return;
    uint4 r0, r1;
    pwf_make_record(F, bin, r0, r1);
    F.bin_run[2 * bin] = r0;
    F.bin_run[2 * bin + 1] = r1;
}

// ---- ONE binning pass (pw_band_bins_kernel) instead of  memset + pw_span_bin_kernel + pw_bin_runs_kernel  ----------------
// A CTA owns a BAND of map rows (as many as fit PWB_NB_MAX bins: 17 rows of a 4K map) and keeps the band's bins — counter,
// whole-bin id, eight entries each — in SHARED memory: the entry slots come from shared-memory atomics (no round trip to
// L2 per entry, the latency that bounded pw_span_bin_kernel), and the run records are built from shared memory straight
// into F.bin_run.  The global counter / entry arrays, their memset and the run pass's read-back are gone.
//
//   A  scan the triangles' row ranges (8 bytes each) for (triangle, row) pairs that can reach the band,
//   B  evaluate those pairs exactly like pw_span_bin_kernel (eight lanes per triangle), clip the interval to the band, bin it,
//   C  one run record per bin (pwf_make_record_from).
//
// Which pairs "can reach the band": the reference's relative index is  (y - yOff) * oW + X  with X = Math.round(x limit)
// NOT offset by xOff, so row y of a triangle lands in map row  r + d,  r = y - yOff,  d = floor(X / oW)  (or, for a negative
// index, oH rows further down: TypedArray.fill counts those from the end).  For X in [xOff - 1, xOff + oW + 1] — one more
// than the window the points' extrema define — d lies in [dlo, dhi] below, and the band evaluates the rows r with
// r + d or r + oH + d inside it.  That assumption is CHECKED, not trusted: every evaluated pair whose interval leaves the
// rows it was assumed to reach flags the frame for the general path, and the (few, as a rule none) rows of triangles no
// band would evaluate at all are evaluated by band 0 for that check alone.  A pair is evaluated by every band it can reach
// ((rows + dhi - dlo) / rows times on average: 19 / 17 for a 4K frame at xOff = 0).
#ifndef HG_PWB_THREADS
#define HG_PWB_THREADS 256
#endif
#ifndef HG_PWB_MINB
#define HG_PWB_MINB 3
#endif
#ifndef HG_PWB_NB
#define HG_PWB_NB 1024
#endif
constexpr int PWB_THREADS = HG_PWB_THREADS;
constexpr int PWB_NB_MAX = HG_PWB_NB;     // bins of one band (40 bytes each)
constexpr int PWB_WORK_CAP = 2048;   // (triangle, candidate row range) segments waiting for evaluation
constexpr int PWB_TRI_STEP = PWB_THREADS;    // triangles scanned between two looks at the worklist (at most four segments each)
constexpr int PWB_LPS = 8;           // lanes per segment
constexpr int PWB_MAX_ROWS = 64;     // rows per band at most
constexpr int PWB_CHECK_ROWS = 64;   // rows of one triangle outside every band's reach that band 0 still evaluates
constexpr size_t PWB_SMEM = (size_t)PWB_NB_MAX * (8 + 4 * PW_BIN_CAP) + (size_t)PWB_WORK_CAP * 4 + 16;

__host__ __device__ inline int pwb_band_rows(int bins_x)
{
    const int r = PWB_NB_MAX / (bins_x > 0 ? bins_x : 1);
    return r > PWB_MAX_ROWS ? PWB_MAX_ROWS : (r < 1 ? 1 : r);
}
__host__ __device__ inline int pwb_floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }   // b > 0

__global__ void __launch_bounds__(PWB_THREADS, HG_PWB_MINB) pw_band_bins_kernel(const FusedFrame *frames)
{
    extern __shared__ uint4 pwb_smem[];
    const FusedFrame &G = frames[blockIdx.y];
    const int oW = G.oW, oH = G.oH;
    if (oW <= 0 || oH <= 0) return;
    const int bins_x = G.bins_x;
    if (bins_x > PWB_NB_MAX) {   // the host does not launch this kernel for such frames
        if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(G.status, 1);
        return;
    }
    const int BR = pwb_band_rows(bins_x), R0 = (int)blockIdx.x * BR;
    if (R0 >= oH) return;
    const int rows = min(BR, oH - R0), nb = rows * bins_x, tid = (int)threadIdx.x;
    unsigned *s_ent = reinterpret_cast<unsigned *>(pwb_smem);
    unsigned *s_cnt = s_ent + PWB_NB_MAX * PW_BIN_CAP, *s_full = s_cnt + PWB_NB_MAX, *s_work = s_full + PWB_NB_MAX;
    unsigned *s_n = s_work + PWB_WORK_CAP;
    for (int i = tid; i < nb; i += PWB_THREADS) {
        s_cnt[i] = 0u;
        s_full[i] = 0u;
    }
    const int xOff = G.xOff, yOff = G.yOff, T = G.n_tris;
    if (tid == 0) {
        *s_n = 0u;
        // one thread divides for the CTA (two signed divisions per thread were 4 % of the pass's instructions)
        s_n[1] = (unsigned)pwb_floor_div(xOff - 1, oW);
        s_n[2] = (unsigned)pwb_floor_div(xOff + oW, oW);
    }
    __syncthreads();
    int *status = G.status;
    const TriRec *rec = G.rec;
    const int2 *yr = G.yr;
    const int dlo = (int)s_n[1], dhi = (int)s_n[2];
    // candidate rows (inclusive, in y): A = rows that reach the band directly, B = through an index counted from the end
    int ya0 = yOff + R0 - dhi, ya1 = yOff + R0 + rows - 1 - dlo, yb0 = ya0 - oH, yb1 = ya1 - oH;
    if (yb1 >= ya0 - 1) {   // the two ranges meet (maps of a few rows): one range, no pair evaluated twice
        ya0 = yb0;
        yb1 = yb0 - 1;
    }
    const bool check_band = blockIdx.x == 0;
    const int yc = yOff - oH - dhi;   // rows y < yc  and ...
    const int yd = yOff + oH - dlo;   // ... rows y >= yd reach no band under the assumption: band 0 verifies that
    const unsigned band_k0 = (unsigned)R0 * (unsigned)oW, band_k1 = (unsigned)(R0 + rows) * (unsigned)oW;
    const unsigned len = (unsigned)oW * (unsigned)oH;
    const double mw = (double)oW, yoff = (double)yOff, dlen = (double)len;

    for (int tb = 0; tb < T; tb += PWB_TRI_STEP) {
        // ---- A: segments of the next PWB_TRI_STEP triangles
#pragma unroll
        for (int u = 0; u < PWB_TRI_STEP / PWB_THREADS; ++u) {
            const int t = tb + u * PWB_THREADS + tid;
            if (t >= T) break;
            const int2 b = __ldg(yr + t);   // rows [b.x, b.y)
            if (max(b.x, ya0) <= min(b.y - 1, ya1)) s_work[atomicAdd(s_n, 1u)] = (unsigned)t;
            if (max(b.x, yb0) <= min(b.y - 1, yb1)) s_work[atomicAdd(s_n, 1u)] = (unsigned)t | (1u << 17);
            if (check_band) {
                const int nc = min(b.y, yc) - b.x, nd = b.y - max(b.x, yd);
                if (nc > PWB_CHECK_ROWS || nd > PWB_CHECK_ROWS) atomicOr(status, 1);   // a window far smaller than the mesh
                else {
                    if (nc > 0) s_work[atomicAdd(s_n, 1u)] = (unsigned)t | (2u << 17);
                    if (nd > 0) s_work[atomicAdd(s_n, 1u)] = (unsigned)t | (3u << 17);
                }
            }
        }
        __syncthreads();
        const int n_work = (int)*s_n;
        __syncthreads();   // everyone has read the count before the next step appends to it
        if (n_work <= PWB_WORK_CAP - 4 * PWB_TRI_STEP && tb + PWB_TRI_STEP < T) continue;   // room for another step (CTA-uniform)
        // ---- B: evaluate, `lps` lanes per segment (8 .. 32: as many as keep the CTA's threads busy with the segments at hand)
        int lps = PWB_LPS;
        while (lps < 32 && n_work * lps < PWB_THREADS) lps <<= 1;
        for (int sg = tid / lps; sg < n_work; sg += PWB_THREADS / lps) {
            const unsigned w = s_work[sg];
            const int t = (int)(w & 0x1FFFFu), which = (int)(w >> 17);
            const int2 b = __ldg(yr + t);
            int lo, hi;
            if (which == 0) { lo = max(b.x, ya0); hi = min(b.y - 1, ya1); }
            else if (which == 1) { lo = max(b.x, yb0); hi = min(b.y - 1, yb1); }
            else if (which == 2) { lo = b.x; hi = min(b.y, yc) - 1; }
            else { lo = max(b.x, yd); hi = b.y - 1; }
            TriRec r;
            {
                const TriRec *__restrict__ g = rec + t;
#pragma unroll
                for (int e = 0; e < 3; ++e) { r.m[e] = __ldg(&g->m[e]); r.b[e] = __ldg(&g->b[e]); r.lo[e] = __ldg(&g->lo[e]); r.hi[e] = __ldg(&g->hi[e]); }
            }
            const unsigned tt = (unsigned)t << 14;
            for (int yy = lo + (tid & (lps - 1)); yy <= hi; yy += lps) {
                const double y = (double)yy;
                double xo, xd;
                predict_x_limits(r, y, xo, xd);
                const double rowbase = __dmul_rn(__dsub_rn(y, yoff), mw);
                const double rel0 = __dadd_rn(rowbase, js_round(xo)), rel1 = __dadd_rn(rowbase, js_round(xd));
                unsigned k0, k1;
                if (rel0 >= 0.0 && rel0 < dlen && rel1 >= 0.0) {
                    k0 = (unsigned)(int)rel0;
                    k1 = rel1 < dlen ? (unsigned)(int)rel1 : len;
                } else {
                    k0 = (unsigned)js_fill_bound(rel0, (long long)len);
                    k1 = (unsigned)js_fill_bound(rel1, (long long)len);
                }
                if (k0 >= k1) continue;
                // the interval must stay inside the rows this pair was assumed to reach (see above)
                const int rr = yy - yOff, ra = (int)(k0 / (unsigned)oW), rz = (int)((k1 - 1u) / (unsigned)oW);
                const bool direct = ra >= rr + dlo && rz <= rr + dhi, wrapped = ra >= rr + oH + dlo && rz <= rr + oH + dhi;
                if (!(direct || wrapped)) {
                    atomicOr(status, 1);
                    continue;
                }
                unsigned a = max(k0, band_k0);
                const unsigned z = min(k1, band_k1);
                if (a >= z) continue;
                unsigned row = a / (unsigned)oW, c0 = a - row * (unsigned)oW;
                while (a < z) {
                    const unsigned room = (unsigned)oW - c0, want = z - a, n = want < room ? want : room;
                    const unsigned c1 = c0 + n;
                    const unsigned base = (row - (unsigned)R0) * (unsigned)bins_x;
                    const unsigned b_first = c0 / PW_BIN_W, b_last = (c1 - 1u) / PW_BIN_W;
                    const unsigned blo = c0 - b_first * PW_BIN_W, bhi = c1 - b_last * PW_BIN_W;
                    unsigned e_bin[2], e_val[2];
                    int ne = 0;
                    if (b_first == b_last) {
                        if (blo == 0u && bhi == (unsigned)PW_BIN_W) atomicMax(s_full + base + b_first, (unsigned)t + 1u);
                        else { e_bin[ne] = base + b_first; e_val[ne++] = tt | (bhi << 7) | blo; }
                    } else {
                        if (blo == 0u) atomicMax(s_full + base + b_first, (unsigned)t + 1u);
                        else { e_bin[ne] = base + b_first; e_val[ne++] = tt | ((unsigned)PW_BIN_W << 7) | blo; }
                        for (unsigned bb = b_first + 1u; bb < b_last; ++bb) atomicMax(s_full + base + bb, (unsigned)t + 1u);
                        if (bhi == (unsigned)PW_BIN_W) atomicMax(s_full + base + b_last, (unsigned)t + 1u);
                        else { e_bin[ne] = base + b_last; e_val[ne++] = tt | (bhi << 7); }
                    }
                    for (int e = 0; e < ne; ++e) {
                        const unsigned slot = atomicAdd(s_cnt + e_bin[e], 1u);
                        if (slot < (unsigned)PW_BIN_CAP) s_ent[e_bin[e] * PW_BIN_CAP + slot] = e_val[e];
                        else atomicOr(status, 1);
                    }
                    a += n;
                    ++row;
                    c0 = 0u;
                }
            }
        }
        __syncthreads();
        if (tid == 0) *s_n = 0u;
        __syncthreads();
    }

    // ---- C: the run records of the band.  Bins without an entry, or with one entry and nothing behind it (most bins of a
    // coarse mesh), are encoded on the spot; the others are collected and resolved in a second, dense sweep — otherwise
    // every warp walks the general path for the few of its lanes that need it.
    uint4 *out = G.bin_run + 2 * (size_t)R0 * (size_t)bins_x;
    for (int i = tid; i < nb; i += PWB_THREADS) {
        const unsigned cnt = s_cnt[i];
        const int full_raw = (int)s_full[i] - 1;
        if (cnt == 0u || (cnt == 1u && full_raw < 0)) {
            uint4 r0, r1;
            pwf_make_record_from(s_ent + (size_t)i * PW_BIN_CAP, cnt, full_raw, T, status, r0, r1);
            out[2 * i] = r0;
            out[2 * i + 1] = r1;
        } else {
            s_work[atomicAdd(s_n, 1u)] = (unsigned)i;   // nb <= PWB_NB_MAX <= PWB_WORK_CAP
        }
    }
    __syncthreads();
    const int n_gen = (int)*s_n;
    for (int j = tid; j < n_gen; j += PWB_THREADS) {
        const int i = (int)s_work[j];
        uint4 r0, r1;
        pwf_make_record_from(s_ent + (size_t)i * PW_BIN_CAP, s_cnt[i], (int)s_full[i] - 1, T, status, r0, r1);
        out[2 * i] = r0;
        out[2 * i + 1] = r1;
    }
}

// id of run r (0..7) from the packed int16 ids
__device__ __forceinline__ int pwf_run_id(const uint4 &ids, unsigned r)
{
    const unsigned lo = (r & 2u) ? ids.y : ids.x, hi = (r & 2u) ? ids.w : ids.z;
    const unsigned w = (r & 4u) ? hi : lo;
    return (int)(short)(unsigned short)((r & 1u) ? (w >> 16) : w);
}

// CTAs per SM the pixel kernels are compiled for.  4 (128 registers, 132 KB of the SM's shared memory, the rest L1) against
// 5 (96 registers): the same on the 4K configs (0.610 / 0.452), +3.5 % on the 1080p video stream (pixel kernel 0.564 ->
// 0.584, whole step 0.447 -> 0.460); 3 loses 12-15 % everywhere, 6 lost 20 %.
#ifndef HG_PWF_MINB
#define HG_PWF_MINB 4
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

// the f32-rounded inverse matrix of triangle t: one 32-byte record (six floats + padding = exactly one sector), widened
// to double once, when a quad moves to another triangle — the pixel loop has no conversions
__device__ __forceinline__ void pwf_load_matrix(const float *inv, int t, double (&m)[6])
{
    const float4 *p = reinterpret_cast<const float4 *>(inv + 8 * (size_t)t);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    m[0] = (double)a.x; m[1] = (double)a.y; m[2] = (double)a.z; m[3] = (double)a.w; m[4] = (double)b.x; m[5] = (double)b.y;
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
//     gathers as asynchronous copies into a shared-memory ring (see pwf_issue) and the CTA's run records staged in
//     shared memory once: with ~30 % fewer instructions than the first generation the loop was latency-bound (ncu:
//     issue-active 47 %) until no register load but the matrix fetch was left in it.
constexpr int PWF_QCAP = 1024;
  // per warp: 16 row groups x 32 lanes x 2 quads, the most a CTA can ever queue

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
    const float *inv;
    unsigned W, npx_src, W2, H2, Wi, Hi, nkflat;   // nkflat = -(HG_HI_ZERO >> 1) * (W + 1): flat = (hy >> 1) * W + (hx >> 1) + nkflat
    int oW, oH, yOff, base0;
    double xs[2][4];
    unsigned vmask[2];
    unsigned long long below[2], inner[2];
    bool prev_bin, has_bin, src_aligned, warp_prev;
    int lane;
    unsigned lt_mask;
    // pipeline state
    int t0[2];           // triangle of each quad's first pixel (S1 -> S2)
    int tm[2];           // triangle whose matrix mq[q] holds
    float4 mqa[2], mqb[2];  // its inverse matrix as loaded (S1 -> S2): (m0..m3), (m4, m5, -, -); widened where it is used
    int qn;              // warp-uniform: entries in this warp's queue
};

// The gathers do not land in registers: every pixel is an ASYNCHRONOUS 4-byte copy global -> shared (cp.async, SASS
// LDGSTS) into the thread's own slot of a ring of PWF_NST row groups, committed as one group per row group and awaited
// (cp.async.wait_group) only when the row group is stored, PWF_DEPTH iterations later.  Why: ptxas tracks every LDG of
// this loop with ONE scoreboard barrier (all gathers and the matrix loads: wr=5 in the SASS), so the first use of any
// loaded register — the next matrix, a pixel to store — waited for ALL loads in flight, the just-issued gathers included
// (ncu: 35 % of the stall samples on the instruction in front of the matrix loads).  Asynchronous copies are not
// scoreboarded at all: PWF_DEPTH row groups of gathers (24 per thread) are in flight behind the arithmetic, and the
// matrix loads are the only register loads left in the loop.  A pixel outside the window is a copy of zero source
// bytes: the hardware fills the slot with zeros (H.js:1047: the output stays transparent).
#ifndef HG_PWF_DEPTH
#define HG_PWF_DEPTH 3
#endif
constexpr int PWF_DEPTH = HG_PWF_DEPTH;   // row groups of gathers in flight
constexpr int PWF_NST = PWF_DEPTH + 1;  // ring slots per thread (the slot being stored is not the one being filled)
// Ring layout: 32-bit words [slot][quad][pixel k][thread] — the 32 lanes of one copy instruction write 32 consecutive
// words (no bank conflict; the [thread][k] layout that a 16-byte read would like is a 4-way conflict for every copy).
constexpr uint32_t PWF_KSTEP = 4u * 128u;   // bytes between pixel k and k+1 of a thread's quad (PWF_THREADS words)

__device__ __forceinline__ void pwf_copy4(uint32_t smem_dst, const uint32_t *src, bool take)
{
    const int n = take ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_dst), "l"(src), "r"(n) : "memory");
}

// S2: coordinates (H.js:1046), window test and Math.round (H.js:1047-1048), flat gather (H.js:1049-1052) of one row group
template <bool ZERO_OFF>
__device__ __forceinline__ void pwf_issue(PwfCtx<ZERO_OFF> &C, const FusedFrame &F, int g, uint32_t slot)
{
    const int row = C.base0 + g * PWF_GROUP_ROWS;
    const double y = (double)(C.yOff + row);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const uint32_t dst = slot + 4u * (uint32_t)(q * 4 * PWF_THREADS);   // word [q][k][thread]: pixel k at dst + k * PWF_KSTEP
        const bool live = (C.t0[q] >= 0) && (row < C.oH) && (C.vmask[q] != 0u);
        const double m0 = (double)C.mqa[q].x, m1 = (double)C.mqa[q].y, m2 = (double)C.mqa[q].z, m3 = (double)C.mqa[q].w,
                     m4 = (double)C.mqb[q].x, m5 = (double)C.mqb[q].y;
        const double r0 = __dmul_rn(m2, y), r1 = __dmul_rn(m3, y);
        if (ZERO_OFF) {
            unsigned hx[4], hy[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                hx[k] = (unsigned)__double2hiint(__fma_rd(affine_coord_exact(m0, C.xs[q][k], r0, m4), 2.0, HG_MAGIC + 1.0));
                hy[k] = (unsigned)__double2hiint(__fma_rd(affine_coord_exact(m1, C.xs[q][k], r1, m5), 2.0, HG_MAGIC + 1.0));
            }
            unsigned flat[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) flat[k] = (hy[k] >> 1) * C.W + ((hx[k] >> 1) + C.nkflat);   // (a >> 1) + b is one LEA.HI
            const unsigned cz = (unsigned)(HG_HI_ZERO + 2);
            const bool ends_inside = live && ((hx[0] - cz) < C.Wi) & ((hy[0] - cz) < C.Hi) & ((hx[3] - cz) < C.Wi) & ((hy[3] - cz) < C.Hi);
            if (__all_sync(0xffffffffu, ends_inside)) {
#pragma unroll
                for (int k = 0; k < 4; ++k) pwf_copy4(dst + PWF_KSTEP * k, C.src + flat[k], true);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool in = live & ((hx[k] - (unsigned)(HG_HI_ZERO + 1)) < C.W2) & ((hy[k] - (unsigned)(HG_HI_ZERO + 1)) < C.H2) &
                                    (flat[k] < C.npx_src);
                    pwf_copy4(dst + PWF_KSTEP * k, C.src + (in ? flat[k] : 0u), in);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                unsigned f = HG_OUTSIDE;
                if (live)
                    f = pwf_decode<false>(affine_coord_exact(m0, C.xs[q][k], r0, m4), affine_coord_exact(m1, C.xs[q][k], r1, m5), F, C.npx_src);
                pwf_copy4(dst + PWF_KSTEP * k, C.src + (f != HG_OUTSIDE ? f : 0u), f != HG_OUTSIDE);
            }
        }
    }
}

// S1: triangle of each quad's first pixel from the run record of row group g, its matrix; note cut quads in the queue
template <bool ZERO_OFF>
__device__ __forceinline__ void pwf_resolve(PwfCtx<ZERO_OFF> &C, int g, unsigned short *wq, const uint4 *rec)
{
    const bool row_ok = C.base0 + g * PWF_GROUP_ROWS < C.oH;
    const uint4 be0 = rec[0], be1 = rec[1];   // shared memory: the record of this thread's row in group g
    // ONE RUN over the whole bin in all four rows of the warp (most row groups of a coarse mesh): both quads belong to run 0,
    // nothing cuts them — no popcounts, no id permute, no cut test; only a first quad that reaches back into the previous
    // bin (frames of unaligned width) is still queued
    if (__all_sync(0xffffffffu, (be0.x == 1u) & (be0.y == 0u))) {
        const int t = (int)(short)(unsigned short)(be1.x & 0xFFFFu);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (t >= 0 && t != C.tm[q]) {
                const float4 *pm = reinterpret_cast<const float4 *>(C.inv + 8 * (size_t)t);
                C.mqa[q] = __ldg(pm);
                C.mqb[q] = __ldg(pm + 1);
                C.tm[q] = t;
            }
            C.t0[q] = t;
        }
        if (C.warp_prev) {   // warp-uniform: some lane of the warp owns a quad that starts in the previous bin
            const bool cut = row_ok && (C.vmask[0] != 0u) && C.prev_bin;
            const unsigned bal = __ballot_sync(0xffffffffu, cut);
            if (cut) wq[C.qn + __popc(bal & C.lt_mask)] = (unsigned short)((g << 6) | (C.lane << 1));
            C.qn += __popc(bal);
        }
        return;
    }
    const unsigned long long m64 = ((unsigned long long)be0.y << 32) | (unsigned long long)be0.x;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const unsigned r = (unsigned)__popcll(m64 & C.below[q]) - 1u;
        const int t = pwf_run_id2(be1, r);
        // rows of one triangle follow each other: the matrix is fetched only when the quad has moved to another triangle
        // (the registers still hold the last triangle's; "no triangle" does not disturb them)
        if (t >= 0 && t != C.tm[q]) {
            const float4 *pm = reinterpret_cast<const float4 *>(C.inv + 8 * (size_t)t);
            C.mqa[q] = __ldg(pm);       // no use of the loaded values in this stage: the next stage widens them, one
            C.mqb[q] = __ldg(pm + 1);   // iteration later, when they have long arrived
            C.tm[q] = t;
        }
        C.t0[q] = t;
        const bool cut = row_ok && (C.vmask[q] != 0u) && (((m64 & C.inner[q]) != 0ull) || (q == 0 && C.prev_bin));
        const unsigned bal = __ballot_sync(0xffffffffu, cut);
        if (cut) wq[C.qn + __popc(bal & C.lt_mask)] = (unsigned short)((g << 6) | (C.lane << 1) | q);
        C.qn += __popc(bal);
    }
}

// S3: stores of one row group: its copies have landed (the caller waited for its commit group), each quad is one
// 16-byte shared-memory read and one 128-bit store
template <bool ZERO_OFF>
__device__ __forceinline__ void pwf_retire(const PwfCtx<ZERO_OFF> &C, int g, uint32_t *p_out, const uint32_t *slot)
{
    if (C.base0 + g * PWF_GROUP_ROWS >= C.oH) return;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        uint32_t *dst = p_out + 32 * q;
        const uint32_t *w = slot + q * 4 * PWF_THREADS;
        const uint32_t px[4] = {w[0], w[PWF_THREADS], w[2 * PWF_THREADS], w[3 * PWF_THREADS]};
        if (C.vmask[q] == 0xFu) {
            *reinterpret_cast<uint4 *>(dst) = make_uint4(px[0], px[1], px[2], px[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (C.vmask[q] & (1u << k)) dst[k] = px[k];
        }
    }
}

template <bool ZERO_OFF>
__device__ __forceinline__ void pwf_body2(const FusedFrame &F, int niter, int tile_x, int row0, unsigned short *wq, uint4 (*s_rec)[2],
                                          uint32_t (*s_px)[2][4][PWF_THREADS], bool records_inline)
{
    const int warp_id = threadIdx.x >> 5;
    const int tx = threadIdx.x & (PWF_TX - 1), ty = threadIdx.x / PWF_TX;
    PwfCtx<ZERO_OFF> C;
    C.lane = threadIdx.x & 31;
    C.lt_mask = (1u << C.lane) - 1u;
    C.src = F.src;
    C.src_aligned = ((unsigned long long)(uintptr_t)F.src & 15ull) == 0ull;
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
    C.nkflat = 0u - (unsigned)(HG_HI_ZERO >> 1) * (C.W + 1u);
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
    C.warp_prev = __any_sync(0xffffffffu, C.prev_bin);
    C.t0[0] = C.t0[1] = -1;
    C.tm[0] = C.tm[1] = -1;
    C.mqa[0] = C.mqa[1] = C.mqb[0] = C.mqb[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    C.qn = 0;

    // S0, once per CTA: the run records of all its rows (16 * ngroups records of 32 bytes, one 64-column bin each) go to
    // shared memory through cp.async — the dependent chain record -> triangle id -> matrix then starts from a 30-cycle
    // shared-memory read instead of a DRAM / L2 round trip in every iteration (ncu of the register-staged variant: 25 % of
    // the stall samples on the first use of the record)
    {
        const int nrec = PWF_GROUP_ROWS * ngroups;
        for (int i = threadIdx.x; i < 2 * nrec; i += PWF_THREADS) {
            const int rr = i >> 1, h = i & 1, row = row0 + rr;
            uint4 *dst = &s_rec[rr][h];
            if (C.has_bin && row < oH) {
                if (records_inline) {
                    // no run-record pass ran for this launch: resolve the bin's span entries here (one thread per record;
                    // the h == 1 thread of the pair idles)
                    if (h == 0) pwf_make_record(F, (size_t)row * F.bins_x + tile_x, s_rec[rr][0], s_rec[rr][1]);
                    continue;
                }
                const uint4 *srcp = F.bin_run + 2 * ((size_t)row * F.bins_x + tile_x) + h;
                const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(srcp) : "memory");
            } else {
                *dst = h ? make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu) : make_uint4(1u, 0u, 0u, 0u);
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
    }

    uint32_t *p_out = F.out + ((long long)C.base0 * oW + xx00);
    const long long out_step = (long long)PWF_GROUP_ROWS * oW;

    // THREE STAGES; iteration `it`:  S3 store row group it-1-DEPTH | S2 coordinates + asynchronous gathers of it-1 |
    // S1 ids + matrix loads of it.  Every iteration commits exactly one copy group (an empty one when it issues nothing),
    // so "all but the newest DEPTH-1 groups have landed" is the condition for storing row group it-1-DEPTH.
    const uint32_t px_base = (uint32_t)__cvta_generic_to_shared(&s_px[0][0][0][threadIdx.x]);
    constexpr uint32_t px_stage_bytes = 2u * 4u * PWF_THREADS * 4u;
#pragma unroll 1
    for (int it = 0; it < ngroups + 1 + PWF_DEPTH; ++it) {
        if (it >= 1 + PWF_DEPTH) {
            asm volatile("cp.async.wait_group %0;" ::"n"(PWF_DEPTH - 1) : "memory");
            const int g = it - 1 - PWF_DEPTH;
            pwf_retire(C, g, p_out, &s_px[(unsigned)g % (unsigned)PWF_NST][0][0][threadIdx.x]);
            p_out += out_step;
        }
        if (it >= 1 && it - 1 < ngroups) pwf_issue(C, F, it - 1, px_base + ((uint32_t)(it - 1) % (uint32_t)PWF_NST) * px_stage_bytes);
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (it < ngroups) pwf_resolve(C, it, wq, s_rec[it * PWF_GROUP_ROWS + ty]);
    }

    // ---- the queued quads again, one QUAD per lane and TWO quads per lane and pass (H.js:1044-1052 verbatim, every pixel
    // with the triangle of its own run).  The loop above has already stored the quad computed with ONE triangle — the run at
    // its first column inside this bin — so only the pixels that belong to ANOTHER run (or to none) are computed again and
    // stored over the provisional ones, one 32-bit store each: for the usual quad that one span boundary cuts that is half
    // of its arithmetic, one matrix instead of two and two gathers instead of four.  Written in phases over both quads so
    // that the dependent round trips overlap: run record(s) -> ids of all pixels -> matrices -> flat indices -> gathers -> stores.
    __syncwarp();  // also orders the provisional stores above before the final ones below
    const int qn = C.qn;
    for (int e0 = 0; e0 < qn; e0 += 64) {
        int row[2], X0[2], tk[2][4], tB[2];
        unsigned redo[2];   // bit k: pixel k is a pixel of the frame whose run is not the one the loop used
        float4 mB[2][2];
        unsigned idx[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int e = e0 + 32 * u + C.lane;
            row[u] = -1;
            X0[u] = 0;
            tB[u] = -1;
            redo[u] = 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) tk[u][k] = -1;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            mB[u][0] = mB[u][1] = z;
            if (e >= qn) continue;
            const unsigned ent = wq[e];
            const int qq = (int)(ent & 1u), sl = (int)((ent >> 1) & 31u), g = (int)(ent >> 6);
            row[u] = row0 + warp_id * 4 + (sl >> 3) + g * PWF_GROUP_ROWS;
            const int ae = (oW & 3) ? (int)(((unsigned)row[u] * (unsigned)oW) & 3u) : 0;
            const int cq = 4 * (sl & 7) + 32 * qq - ae;     // first column of the quad inside this bin (negative: previous bin)
            X0[u] = tile_x * PW_BIN_W + cq;
            // the quad lies in one bin, except a first quad that reaches back into the previous one
            const int binA = X0[u] < 0 ? 0 : (X0[u] >> 6), binB = (X0[u] + 3) >> 6;
            const int rr = row[u] - row0;
            const uint4 t0r = s_rec[rr][0], t1r = s_rec[rr][1];   // this bin's record
            uint4 a0 = t0r, a1 = t1r;
            if (binA != tile_x) {   // the previous bin (a first quad reaching back over the tile's left edge)
                const uint4 *pr = F.bin_run + 2 * ((size_t)row[u] * F.bins_x + binA);
                a0 = __ldg(pr);
                a1 = __ldg(pr + 1);
            }
            uint4 b0 = a0, b1 = a1;
            if (binB != binA && binB < F.bins_x) {   // binB == tile_x
                b0 = t0r;
                b1 = t1r;
            }
            // the run the loop computed the whole quad with (pwf_resolve: the run at column max(cq, 0) of this bin)
            const unsigned long long mt = ((unsigned long long)t0r.y << 32) | (unsigned long long)t0r.x;
            const int t_main = pwf_run_id2(t1r, (unsigned)__popcll(mt & ((2ull << (cq < 0 ? 0 : cq)) - 1ull)) - 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int X = X0[u] + k;
                if (X < 0 || X >= oW) continue;
                const bool inB = (X >> 6) != binA;
                const uint4 &r0 = inB ? b0 : a0, &r1 = inB ? b1 : a1;
                const unsigned long long m64 = ((unsigned long long)r0.y << 32) | (unsigned long long)r0.x;
                const int t = pwf_run_id2(r1, (unsigned)__popcll(m64 & ((2ull << (X & 63)) - 1ull)) - 1u);
                tk[u][k] = t;
                if (t != t_main) {
                    redo[u] |= 1u << k;
                    if (t >= 0 && tB[u] < 0) tB[u] = t;
                }
            }
            if (tB[u] >= 0) {
                const float4 *pm = reinterpret_cast<const float4 *>(F.inv + 8 * (size_t)tB[u]);
                mB[u][0] = __ldg(pm);
                mB[u][1] = __ldg(pm + 1);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const double y = (double)(F.yOff + row[u]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int t = tk[u][k];
                idx[u][k] = HG_OUTSIDE;
                if (t < 0 || !(redo[u] & (1u << k))) continue;
                float4 p0 = mB[u][0], p1 = mB[u][1];
                if (t != tB[u]) {   // a third triangle inside one quad: fetched on the spot
                    const float4 *pm = reinterpret_cast<const float4 *>(F.inv + 8 * (size_t)t);
                    p0 = __ldg(pm);
                    p1 = __ldg(pm + 1);
                }
                const double x = (double)(F.xOff + X0[u] + k);
                idx[u][k] = pwf_decode<ZERO_OFF>(affine_coord_exact((double)p0.x, x, __dmul_rn((double)p0.z, y), (double)p1.x),
                                                 affine_coord_exact((double)p0.y, x, __dmul_rn((double)p0.w, y), (double)p1.y), F, C.npx_src);
            }
        }
        uint32_t v[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int k = 0; k < 4; ++k) v[u][k] = ldg_or_zero(C.src, idx[u][k]);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (redo[u] == 0u) continue;
            uint32_t *dst = F.out + ((long long)row[u] * oW + X0[u]);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (redo[u] & (1u << k)) dst[k] = v[u][k];   // a pixel without a triangle stays transparent: v = 0
        }
    }
}

__global__ void __launch_bounds__(PWF_THREADS, HG_PWF_MINB) pw_warp_fused_kernel(const FusedFrame *frames, int niter, int records_inline)
{
    __shared__ unsigned short s_q[PWF_THREADS / 32][PWF_QCAP];
    const FusedFrame F = frames[blockIdx.y];
    if (F.oW <= 0 || F.oH <= 0) return;
    const int tiles_x = pwf_tiles_x(F.oW);
    const int tile_y = blockIdx.x / tiles_x;
    const int tile_x = blockIdx.x - tile_y * tiles_x;
    const int row0 = tile_y * PWF_GROUP_ROWS * niter;
    if (row0 >= F.oH) return;
    __shared__ uint4 s_rec[PWF_GROUP_ROWS * 16][2];   // run records of the CTA's rows (niter <= 16)
    __shared__ uint32_t s_px[PWF_NST][2][4][PWF_THREADS];   // gathered pixels: a ring of row groups, [quad][pixel][thread]
    unsigned short *wq = s_q[threadIdx.x >> 5];
    // (a frame of unaligned width always has its records in global memory: the host ran the record pass for the launch)
    const bool inl = records_inline != 0 && (F.oW & 3) == 0;
    if (F.minSrcX == 0 && F.minSrcY == 0) pwf_body2<true>(F, niter, tile_x, row0, wq, s_rec, s_px, inl);
    else pwf_body2<false>(F, niter, tile_x, row0, wq, s_rec, s_px, inl);
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
