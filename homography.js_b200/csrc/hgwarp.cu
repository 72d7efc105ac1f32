// hgwarp.cu — context, memory and the extern "C" surface declared in include/hgwarp.h.
// Build (see __graft_entry__.build / homography.js_b200/build.py):
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -shared -Xcompiler -fPIC
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/hgwarp.h"
#include "jsnum.cuh"
#include "piecewise.cuh"
#include "solve.cuh"
#include "warp_geo.cuh"
#include "diag.cuh"
#include "bilinear.cuh"
#include "forward.cuh"
#include "piecewise_fused.cuh"
#include "stream.cuh"
#include "delaunay_host.cuh"
#include "png_host.cuh"
#include "jpeg_host.cuh"

using namespace hg;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct hg_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    uint64_t launches = 0;
    // optional per-kernel timing of the pixel-loop kernels (roofline accounting in bench.py)
    bool prof = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_ev;
    size_t prof_used = 0;

    // image (this._image)
    const uint32_t *img = nullptr;
    DevBuf img_own;
    int W = 0, H = 0;

    // output of the last non-batched warp
    DevBuf out;
    size_t out_bytes = 0;

    // small device scratch: matrices, limits, points (doubles)
    DevBuf scratch;       // 4 KiB
    void *pinned = nullptr;  // 4 KiB pinned host mirror

    // mesh
    DevBuf src_pts, dst_pts, tris, rec, map32, map16, frames, mats, winner;
    DevBuf invd, bin_cnt, bin_ent, bin_run, fstatus, fframes;  // fused piecewise path
    // second set of the fused path's per-chunk scratch: chunk k+1 is binned (stream_pre) while chunk k's pixels are written
    DevBuf rec_b, invd_b, bin_cnt_b, bin_ent_b, bin_run_b, fstatus_b, fframes_b, yr_b;
    DevBuf yr;                                     // row ranges of the triangles (band binning pass)
    cudaStream_t stream_pre = nullptr;             // high-priority lane of the span / run passes
    cudaEvent_t ev_pre[2] = {nullptr, nullptr}, ev_pix[2] = {nullptr, nullptr};
    DevBuf fwd_args, cs_frames, cs_out, sinfo, pts_batch;     // forward batches, checksums, stream bookkeeping, point batches
    bool winner_clean = false;  // every entry of `winner` is -1 (the gather pass hands the plane back clean)
    void *pin_big = nullptr;    // pinned host staging for per-frame status / info read-backs
    cudaEvent_t ev_chunk[2] = {nullptr, nullptr};  // chunk boundaries of hg_warp_piecewise_stream
    cudaStream_t stream2 = nullptr;                // second lane of forward batches (scatter of one sub-batch beside the gather of another)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    size_t pin_big_cap = 0;
    // hg_solve_with_limits runs on a stream and scratch of its own: it depends on nothing queued on the context stream, and
    // the class surface calls it between setImage() (an asynchronous upload) and warp() — the solve's round trip and the
    // host work around it then run beside the upload instead of behind it
    cudaStream_t stream_aux = nullptr;
    DevBuf scratch_aux;   // 4 KiB, same layout as `scratch`
    // hg_pcie_probe keeps its buffers between calls (see there)
    void *probe_h[2] = {nullptr, nullptr}, *probe_d[2] = {nullptr, nullptr};
    cudaStream_t probe_s[2] = {nullptr, nullptr};
    cudaEvent_t probe_e[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t probe_bytes = 0;
    // parameters of the last inverse index map (rebuilt on demand for the aliasing forward read, Q8)
    std::vector<float> last_inv_pts;
    double last_inv_mw = 0, last_inv_yoff = 0;
    long long last_inv_len = -1;
    bool map32_current = false;
    int sampling = 0;       // HG_NEAREST (the reference) | HG_BILINEAR (extension; inverse affine / projective only)
    int force_general = 0;  // diagnostics: 1 = always use the map-based general path
    uint64_t n_fused = 0, n_general = 0;  // inverse piecewise frames finished by each path
    int n_pts = 0, n_tris = 0;
    long long map_len = 0;  // length of the map currently in map32 (for the aliasing forward read)

    // TMA staging of source tiles (warp_geo.cuh): cuTensorMapEncodeTiled from the driver, the tensor maps of the
    // context image (passed to single-frame launches by value) and a device-side cache for batch frames
    void *tm_encode = nullptr;
    int geo_box_bytes = 13824;      // shared memory for one staged source box (HG_GEO_BOX_BYTES): 72 x 48 pixels
    int geo_niter_staged = 2;       // row groups (16 rows) per tile of the staged kernel (HG_GEO_NITER)
    int geo_stages = 3;             // ring depth per CTA (HG_GEO_STAGES)
    int geo_ctas_per_sm = 5;        // persistent CTAs per SM (HG_GEO_CTAS; bounded by registers / shared memory)
    int geo_debug = 0;              // HG_GEO_DEBUG: the staged kernel's producer traces its ring
    bool no_tall = false;           // HG_GEO_NO_TALL: never pick the tall thread layout (A/B runs)
    bool bilinear_v1 = false;       // HG_BILINEAR_V1: first-generation bilinear kernel (A/B runs)
    bool pwf_v1 = false;            // HG_PWF_V1: first-generation fused piecewise pixel kernel (A/B runs)
    bool geo_async = false;         // HG_GEO_ASYNC: affine / projective pixel loop with asynchronous gathers (A/B runs)
    int pw_chunk = 0;               // HG_PW_CHUNK: frames per pipelined chunk of the piecewise batch / stream calls (0: 128; 64 with two lanes)
    int fwd_plane_mb = 48;          // HG_FWD_PLANE_MB: budget of the forward batches' winner planes (kept L2-resident)
    int pw_binning = 0;             // HG_PW_BINNING / hg_debug_piecewise_binning: 0 auto, 1 span + run passes, 2 one band pass
    int pw_lanes = -1;              // HG_PW_LANES: 0 one lane (every chunk's binning passes in front of its pixel kernel), 1 two lanes,
                                    // unset: two lanes for meshes binned by the span + run passes, one for the band pass
    bool pwf_records_inline = false;  // HG_PWF_RECORDS_INLINE: the pixel kernel builds the run records of aligned frames itself (A/B runs)
    CUtensorMap img_tm[GEO_NBOX];
    bool img_tm_ok = false;
    DevBuf tm_dev;                   // TM_CACHE_SLOTS x GEO_NBOX tensor maps
    std::unordered_map<uint64_t, std::vector<std::pair<uint64_t, int>>> tm_index;  // hash -> [(ptr ^ dims, slot)]
    int tm_used = 0;
};

namespace {

// scratch layout (bytes): [0,128) src pts, [128,256) dst pts, [256,320) matrix, [320,352) limits, [384,512) misc
constexpr size_t SC_SRC = 0, SC_DST = 128, SC_MAT = 256, SC_LIM = 320, SC_MISC = 384;

int fail(hg_ctx *c, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    else g_create_error = buf;
    return code;
}

#define CU(c, call)                                                                              \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return fail((c), HG_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                     \
    } while (0)

#define NEED(c, cond, msg)                                          \
    do {                                                            \
        if (!(cond)) return fail((c), HG_ERR_INVALID, "%s", (msg)); \
    } while (0)

int ensure(hg_ctx *c, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return HG_OK;
    CU(c, cudaStreamSynchronize(c->stream));
    if (b.p) CU(c, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(c, HG_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    b.cap = want;
    return HG_OK;
}

int ensure_pinned(hg_ctx *c, size_t bytes)
{
    if (bytes <= c->pin_big_cap) return HG_OK;
    CU(c, cudaStreamSynchronize(c->stream));
    if (c->pin_big) CU(c, cudaFreeHost(c->pin_big));
    c->pin_big = nullptr;
    c->pin_big_cap = 0;
    const size_t want = bytes + bytes / 2 + 4096;
    cudaError_t e = cudaHostAlloc(&c->pin_big, want, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(c, HG_ERR_NOMEM, "cudaHostAlloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    c->pin_big_cap = want;
    return HG_OK;
}

int bind(hg_ctx *c)
{
    CU(c, cudaSetDevice(c->device));
    return HG_OK;
}

#define BIND(c)                                   \
    do {                                          \
        if (!(c)) return HG_ERR_INVALID;          \
        int r_ = bind(c);                         \
        if (r_) return r_;                        \
    } while (0)

#define TRY(expr)                 \
    do {                          \
        int r_ = (expr);          \
        if (r_) return r_;        \
    } while (0)

int check_window(hg_ctx *c, int x_off, int y_off, int o_w, int o_h)
{
    if (o_w < 1 || o_h < 1) return fail(c, HG_ERR_INVALID, "output size %dx%d must be >= 1x1", o_w, o_h);
    if (o_w > 65536 || o_h > 65536 || (long long)o_w * o_h >= (1LL << 31))
        return fail(c, HG_ERR_UNSUPPORTED, "output size %dx%d outside the supported range", o_w, o_h);
    if (x_off > (1 << 18) || x_off < -(1 << 18) || y_off > (1 << 18) || y_off < -(1 << 18))
        return fail(c, HG_ERR_UNSUPPORTED, "output offset (%d,%d) outside the supported range", x_off, y_off);
    return HG_OK;
}

int check_image_dims(hg_ctx *c, int w, int h)
{
    if (w < 1 || h < 1) return fail(c, HG_ERR_INVALID, "image size %dx%d must be >= 1x1", w, h);
    if (w > 65536 || h > 65536 || (long long)w * h >= (1LL << 31))
        return fail(c, HG_ERR_UNSUPPORTED, "image size %dx%d outside the supported range", w, h);
    return HG_OK;
}

// where a non-batched warp writes: caller's device buffer or the context's own
int pick_out(hg_ctx *c, void *out_dev, size_t bytes, uint32_t **dst)
{
    if (out_dev) {
        if (((uintptr_t)out_dev & 15) != 0) return fail(c, HG_ERR_INVALID, "out_dev must be 16-byte aligned");
        *dst = (uint32_t *)out_dev;
    } else {
        TRY(ensure(c, c->out, bytes));
        *dst = (uint32_t *)c->out.p;
        c->out_bytes = bytes;
    }
    return HG_OK;
}

int finish_out(hg_ctx *c, const uint32_t *dst, size_t bytes, uint8_t *out_host)
{
    CU(c, cudaGetLastError());
    if (out_host) {
        CU(c, cudaMemcpyAsync(out_host, dst, bytes, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
    }
    return HG_OK;
}

unsigned grid_for(hg_ctx *c, long long npix, int n_frames)
{
    const long long nquad = (npix + 3) / 4;
    long long blocks = (nquad + 255) / 256;
    // enough resident CTAs to cover HBM latency; whole multiples of the SM count per frame
    long long cap = (long long)c->sm_count * 8;
    if (n_frames > 1) cap = (long long)c->sm_count * 2;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

// events bracketing one pixel-loop kernel (only when hg_profile_enable(ctx, 1))
int prof_begin(hg_ctx *c)
{
    if (!c->prof) return HG_OK;
    if (c->prof_used == c->prof_ev.size()) {
        cudaEvent_t a, b;
        CU(c, cudaEventCreate(&a));
        CU(c, cudaEventCreate(&b));
        c->prof_ev.emplace_back(a, b);
    }
    CU(c, cudaEventRecord(c->prof_ev[c->prof_used].first, c->stream));
    return HG_OK;
}

int prof_end(hg_ctx *c)
{
    if (!c->prof) return HG_OK;
    CU(c, cudaEventRecord(c->prof_ev[c->prof_used].second, c->stream));
    c->prof_used++;
    return HG_OK;
}

constexpr int TM_CACHE_SLOTS = 4096;

typedef CUresult (*tm_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// GEO_NBOX tensor maps over a device image seen as a [H][W] tensor of 32-bit pixels, one per box width
// and height (geo_box_w(i % 8) pixels x 8 or 32 rows, out-of-image elements read as zero).  False when the image
// cannot be described (row pitch or base not 16-byte aligned) — the kernels then gather directly.
bool encode_tmaps(hg_ctx *c, const void *img, int W, int H, CUtensorMap *out)
{
    if (!c->tm_encode || !img || (W & 3) != 0 || ((uintptr_t)img & 15) != 0) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
    const cuuint64_t strides[1] = {(cuuint64_t)W * 4};
    const cuuint32_t estr[2] = {1, 1};
    for (int i = 0; i < GEO_NBOX; ++i) {
        const cuuint32_t box[2] = {(cuuint32_t)geo_box_w(i % GEO_NBOX_W),
                                   (cuuint32_t)(i < GEO_NBOX_W ? GEO_BOX_ROWS : GEO_BOX_ROWS_TALL)};
        const CUresult r = ((tm_encode_fn)c->tm_encode)(&out[i], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void *>(img),
                                                         dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return false;
    }
    return true;
}

// device copy of the tensor maps of a borrowed image (batch frames), cached by (pointer, size); nullptr = no staging
int tmaps_device(hg_ctx *c, const void *img, int W, int H, const CUtensorMap **out)
{
    *out = nullptr;
    if (!c->tm_encode || (W & 3) != 0 || ((uintptr_t)img & 15) != 0) return HG_OK;
    const uint64_t key = (uint64_t)(uintptr_t)img, dims = ((uint64_t)(uint32_t)W << 32) | (uint32_t)H;
    auto &bucket = c->tm_index[key];
    for (auto &e : bucket)
        if (e.first == dims) {
            *out = (const CUtensorMap *)c->tm_dev.p + (size_t)e.second * GEO_NBOX;
            return HG_OK;
        }
    CUtensorMap tm[GEO_NBOX];
    if (!encode_tmaps(c, img, W, H, tm)) return HG_OK;
    if (!c->tm_dev.p) TRY(ensure(c, c->tm_dev, sizeof(CUtensorMap) * GEO_NBOX * (size_t)TM_CACHE_SLOTS));
    if (c->tm_used == TM_CACHE_SLOTS) {
        // cache full: this frame gathers directly (*out stays nullptr).  Slots are never recycled — frames resolved
        // earlier in the same batch still point at them until their launch
        if (bucket.empty()) c->tm_index.erase(key);
        return HG_OK;
    }
    const int slot = c->tm_used++;
    CUtensorMap *dst = (CUtensorMap *)c->tm_dev.p + (size_t)slot * GEO_NBOX;
    // pageable source: staged before the call returns, ordered on the context stream before any launch that uses it
    CU(c, cudaMemcpyAsync(dst, tm, sizeof tm, cudaMemcpyHostToDevice, c->stream));
    bucket.emplace_back(dims, slot);
    *out = dst;
    return HG_OK;
}

// rows per CTA = 64 * niter: long-lived CTAs amortise their start-up and pipeline gathers against arithmetic,
// but the grid must still fill the machine (>= ~6 CTAs per SM in total)
int pick_niter(hg_ctx *c, int max_ow, int max_oh, int n_frames, int ltx = 4)
{
    int niter = ltx == 4 ? 16 : 2;  // tall row groups are 128 rows already
    while (niter > 1 && (long long)geo_tiles_x(max_ow, 1 << ltx) * geo_tiles_y(max_oh, niter, geo_group_rows(ltx)) * n_frames <
                            (long long)c->sm_count * 16)
        niter >>= 1;
    return niter;
}

// Does the inverse map turn the image by about a quarter turn (a step in output x is mostly a step in source y)?  Then
// the "tall" thread layout gathers coalesced (warp_geo.cuh, geo_group_rows).  From the matrix: the derivative of the
// source coordinates along output x at the centre of the window.
bool map_is_rotated(int kind, const double *m, int x_off, int y_off, int o_w, int o_h)
{
    double dsx, dsy;
    if (kind == HG_AFFINE) {
        dsx = m[0];
        dsy = m[1];
    } else {
        const double x = x_off + 0.5 * o_w, y = y_off + 0.5 * o_h;
        const double dn = m[6] * x + m[7] * y + 1.0, nx = m[0] * x + m[1] * y + m[2], ny = m[3] * x + m[4] * y + m[5];
        dsx = m[0] * dn - m[6] * nx;
        dsy = m[3] * dn - m[6] * ny;
    }
    return std::isfinite(dsx) && std::isfinite(dsy) && std::fabs(dsy) > 2.0 * std::fabs(dsx);
}

// the same question from point pairs (the matrix is solved on the device): affine estimate through the first three pairs
// of the inverse map  from[i] -> to[i]
bool points_map_is_rotated(const double *from, const double *to)
{
    const double ax = from[2] - from[0], ay = from[3] - from[1], bx = from[4] - from[0], by = from[5] - from[1];
    const double det = ax * by - ay * bx;
    if (!(std::fabs(det) > 0.0)) return false;
    // d(to)/d(from.x) of the affine map through the three pairs
    const double ux = to[2] - to[0], uy = to[3] - to[1], vx = to[4] - to[0], vy = to[5] - to[1];
    const double m[6] = {(ux * by - vx * ay) / det, (uy * by - vy * ay) / det, 0, 0, 0, 0};
    return map_is_rotated(HG_AFFINE, m, 0, 0, 1, 1);
}

// `staged`: the frames carry tensor maps (P.has_tm / GeoFrame::tm), so CTAs cover few rows and stage their source
// footprint in shared memory; otherwise long-lived CTAs gather directly
int launch_geo(hg_ctx *c, int kind, GeoParams &P, int max_ow, int max_oh, int n_frames, bool staged, cudaStream_t stream,
               bool tall = false)
{
    P.ltx = (tall && !c->no_tall && !staged && c->sampling != HG_BILINEAR) ? 1 : 4;
    P.niter = staged ? c->geo_niter_staged : pick_niter(c, max_ow, max_oh, n_frames, P.ltx);
    if (staged && 32 * P.niter > GEO_QCAP) P.niter = GEO_QCAP / 32;  // per-tile exact queue: one entry per thread and row group
    P.box_bytes = staged ? c->geo_box_bytes : 0;
    P.stages = 0;
    dim3 grid((unsigned)(geo_tiles_x(max_ow, 1 << P.ltx) * geo_tiles_y(max_oh, P.niter, geo_group_rows(P.ltx))), (unsigned)n_frames);
    if (c->sampling == HG_BILINEAR) {
        const long long nq = ((long long)max_ow * max_oh + 3) / 4;
        long long blocks = (nq + 255) / 256;
        // several quads per thread amortise the per-thread set-up (frame and matrix loads, geo_fast_mode) as long as the
        // grid still fills the machine several times over
        int qpt = 4;
        while (qpt > 1 && (blocks / qpt) * n_frames < (long long)c->sm_count * 8) qpt >>= 1;
        blocks = (blocks + qpt - 1) / qpt;
        if (blocks > (long long)c->sm_count * 16) blocks = (long long)c->sm_count * 16;
        dim3 g2((unsigned)blocks, (unsigned)n_frames);
        TRY(prof_begin(c));
        if (c->bilinear_v1) {  // first-generation kernel, kept for A/B runs
            if (kind == HG_AFFINE) warp_inverse_geo_bilinear_kernel<0><<<g2, 256, 0, stream>>>(P);
            else warp_inverse_geo_bilinear_kernel<1><<<g2, 256, 0, stream>>>(P);
        } else {
            if (kind == HG_AFFINE) warp_inverse_geo_bilinear2_kernel<0><<<g2, 256, 0, stream>>>(P);
            else warp_inverse_geo_bilinear2_kernel<1><<<g2, 256, 0, stream>>>(P);
        }
        c->launches++;
        CU(c, cudaGetLastError());
        TRY(prof_end(c));
        return HG_OK;
    }
    const bool timed = stream == c->stream;  // the per-kernel events live on the context stream
    if (timed) TRY(prof_begin(c));
    if (staged) {
        P.stages = c->geo_stages;
        P.tiles_x = geo_tiles_x(max_ow);
        P.tiles_y = geo_tiles_y(max_oh, P.niter);
        P.n_frames = n_frames;
        P.debug = c->geo_debug;  // 1: trace ring entries, 2: no 32-row boxes
        const long long total = (long long)P.tiles_x * P.tiles_y * n_frames;
        long long ctas = (long long)c->sm_count * c->geo_ctas_per_sm;
        if (ctas > total) ctas = total;
        const size_t smem = (size_t)P.stages * (size_t)(GEO_HDR_BYTES + P.box_bytes);
        if (kind == HG_AFFINE) warp_inverse_geo_staged_kernel<0><<<(unsigned)ctas, GEO_STAGED_THREADS, smem, stream>>>(P);
        else warp_inverse_geo_staged_kernel<1><<<(unsigned)ctas, GEO_STAGED_THREADS, smem, stream>>>(P);
    } else if (c->geo_async) {
        if (kind == HG_AFFINE) warp_inverse_geo_async_kernel<0><<<grid, GEO_THREADS, 0, stream>>>(P);
        else warp_inverse_geo_async_kernel<1><<<grid, GEO_THREADS, 0, stream>>>(P);
    } else {
        if (kind == HG_AFFINE && P.ltx == 1) warp_inverse_geo_affine_tall_kernel<<<grid, GEO_THREADS, 0, stream>>>(P);
        else if (kind == HG_AFFINE) warp_inverse_geo_kernel<0><<<grid, GEO_THREADS, 0, stream>>>(P);
        else warp_inverse_geo_kernel<1><<<grid, GEO_THREADS, 0, stream>>>(P);
    }
    c->launches++;
    CU(c, cudaGetLastError());
    if (timed) TRY(prof_end(c));
    return HG_OK;
}

int launch_solve(hg_ctx *c, const SolveArgs &a, cudaStream_t st = nullptr)
{
    solve_kernel<<<(a.n + 63) / 64, 64, 0, st ? st : c->stream>>>(a);
    c->launches++;
    CU(c, cudaGetLastError());
    return HG_OK;
}


int upload_small(hg_ctx *c, size_t off, const void *host, size_t bytes)
{
    // pageable -> device; tiny, goes through the driver's staging buffer, ordered on the stream
    CU(c, cudaMemcpyAsync((char *)c->scratch.p + off, host, bytes, cudaMemcpyHostToDevice, c->stream));
    return HG_OK;
}

int download_small(hg_ctx *c, void *host, size_t off, size_t bytes)
{
    CU(c, cudaMemcpyAsync(host, (char *)c->scratch.p + off, bytes, cudaMemcpyDeviceToHost, c->stream));
    return HG_OK;
}

}  // namespace

extern "C" {

int hg_abi_version(void) { return HGWARP_ABI_VERSION; }

int hg_device_count(int *count)
{
    if (!count) return HG_ERR_INVALID;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        return fail(nullptr, HG_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    }
    *count = n;
    return HG_OK;
}

int hg_ctx_create(int device, hg_ctx **out)
{
    if (!out) return HG_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(nullptr, HG_ERR_CUDA, "no CUDA device available: %s", cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return fail(nullptr, HG_ERR_INVALID, "device %d out of range [0,%d)", device, n);
    hg_ctx *c = new hg_ctx();
    c->device = device;
#define CUC(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            fail(nullptr, HG_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));                 \
            cudaGetLastError();                                                                         \
            hg_ctx_destroy(c); /* releases whatever was created so far */                               \
            return HG_ERR_CUDA;                                                                         \
        }                                                                                               \
    } while (0)
    CUC(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUC(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    CUC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUC(cudaEventCreate(&c->ev0));
    CUC(cudaEventCreate(&c->ev1));
    CUC(cudaMalloc(&c->scratch.p, 4096));
    c->scratch.cap = 4096;
    CUC(cudaMemset(c->scratch.p, 0, 4096));
    CUC(cudaHostAlloc(&c->pinned, 4096, cudaHostAllocDefault));
    {
        // TMA staging (warp_inverse_geo_staged_kernel) is opt-in, HG_GEO_STAGED=1: measured on B200 it is slower than
        // the direct-gather kernel for these nearest-neighbour maps; the tensor-map encoder lives in
        // the driver
        const char *on = getenv("HG_GEO_STAGED");
        cudaDriverEntryPointQueryResult qres;
        void *fn = nullptr;
        if ((on && on[0] == '1') &&
            cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            c->tm_encode = fn;
        cudaGetLastError();
        auto env_int = [](const char *name, int lo, int hi, int &dst) {
            if (const char *v = getenv(name)) {
                const int n = atoi(v);
                if (n >= lo && n <= hi) dst = n;
            }
        };
        env_int("HG_GEO_BOX_BYTES", 4096, 100 * 1024, c->geo_box_bytes);
        c->geo_box_bytes &= ~127;
        env_int("HG_GEO_NITER", 1, 16, c->geo_niter_staged);
        env_int("HG_GEO_STAGES", 2, GEO_MAX_STAGES, c->geo_stages);
        env_int("HG_GEO_CTAS", 1, 8, c->geo_ctas_per_sm);
        env_int("HG_GEO_DEBUG", 0, 2, c->geo_debug);
        c->no_tall = getenv("HG_GEO_NO_TALL") != nullptr;
        c->bilinear_v1 = getenv("HG_BILINEAR_V1") != nullptr;
        c->pwf_v1 = getenv("HG_PWF_V1") != nullptr;
        c->geo_async = getenv("HG_GEO_ASYNC") != nullptr;
        c->pwf_records_inline = getenv("HG_PWF_RECORDS_INLINE") != nullptr;
        env_int("HG_PW_CHUNK", 1, 1024, c->pw_chunk);
        env_int("HG_PW_LANES", 0, 1, c->pw_lanes);
        env_int("HG_PW_BINNING", 0, 2, c->pw_binning);
        env_int("HG_FWD_PLANE_MB", 1, 4096, c->fwd_plane_mb);
        CUC(cudaFuncSetAttribute(pw_band_bins_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PWB_SMEM));
        // the ring must fit a CTA's shared memory: shrink the depth (the CTA count follows from the occupancy below)
        const size_t cta_max = prop.sharedMemPerBlockOptin;
        auto ring = [&]() { return (size_t)c->geo_stages * (size_t)(GEO_HDR_BYTES + c->geo_box_bytes); };
        while (c->geo_stages > 2 && ring() + 4096 > cta_max) c->geo_stages--;
        if (ring() + 4096 > cta_max) c->tm_encode = nullptr;  // box does not fit at all: direct kernel only
        CUC(cudaFuncSetAttribute(warp_inverse_geo_staged_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring()));
        CUC(cudaFuncSetAttribute(warp_inverse_geo_staged_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring()));
        // persistent grid: never more CTAs per SM than can be resident (registers / shared memory)
        int occ0 = 0, occ1 = 0;
        CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ0, warp_inverse_geo_staged_kernel<0>, GEO_STAGED_THREADS, ring()));
        CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, warp_inverse_geo_staged_kernel<1>, GEO_STAGED_THREADS, ring()));
        const int occ = occ0 < occ1 ? occ0 : occ1;
        if (occ < 1) c->tm_encode = nullptr;
        else if (c->geo_ctas_per_sm > occ) c->geo_ctas_per_sm = occ;
        if (getenv("HG_GEO_VERBOSE"))
            fprintf(stderr, "hgwarp: staged kernel: box %d B, %d stages, niter %d, %d CTAs/SM (occupancy %d/%d), tma %s\n",
                    c->geo_box_bytes, c->geo_stages, c->geo_niter_staged, c->geo_ctas_per_sm, occ0, occ1,
                    c->tm_encode ? "on" : "off");
    }
#undef CUC
    *out = c;
    return HG_OK;
}

static void hg_pcie_probe_release(hg_ctx *c)
{
    for (auto &ev : c->probe_e) { if (ev) cudaEventDestroy(ev); ev = nullptr; }
    for (auto &st : c->probe_s) { if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); } st = nullptr; }
    for (auto &p : c->probe_d) { if (p) cudaFree(p); p = nullptr; }
    for (auto &p : c->probe_h) { if (p) cudaFreeHost(p); p = nullptr; }
    c->probe_bytes = 0;
}

int hg_ctx_destroy(hg_ctx *c)
{
    if (!c) return HG_ERR_INVALID;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    hg_pcie_probe_release(c);
    if (c->stream_aux) {
        cudaStreamSynchronize(c->stream_aux);
        cudaStreamDestroy(c->stream_aux);
    }
    if (c->scratch_aux.p) cudaFree(c->scratch_aux.p);
    DevBuf *bufs[] = {&c->img_own, &c->out, &c->scratch, &c->src_pts, &c->dst_pts, &c->tris,
                      &c->rec, &c->map32, &c->map16, &c->frames, &c->mats, &c->winner,
                      &c->invd, &c->bin_cnt, &c->bin_ent, &c->bin_run, &c->fstatus, &c->fframes, &c->tm_dev,
                      &c->rec_b, &c->invd_b, &c->bin_cnt_b, &c->bin_ent_b, &c->bin_run_b, &c->fstatus_b, &c->fframes_b,
                      &c->yr, &c->yr_b,
                      &c->fwd_args, &c->cs_frames, &c->cs_out, &c->sinfo, &c->pts_batch};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->pin_big) cudaFreeHost(c->pin_big);
    for (cudaEvent_t e : c->ev_chunk)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_pre)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_pix)
        if (e) cudaEventDestroy(e);
    if (c->stream_pre) {
        cudaStreamSynchronize(c->stream_pre);
        cudaStreamDestroy(c->stream_pre);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->stream2) {
        cudaStreamSynchronize(c->stream2);
        cudaStreamDestroy(c->stream2);
    }
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (auto &pr : c->prof_ev) {
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return HG_OK;
}

const char *hg_last_error(hg_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int hg_ctx_synchronize(hg_ctx *c)
{
    BIND(c);
    CU(c, cudaStreamSynchronize(c->stream));
    return HG_OK;
}

int hg_ctx_stream(hg_ctx *c, void **s)
{
    if (!c || !s) return HG_ERR_INVALID;
    *s = (void *)c->stream;
    return HG_OK;
}

int hg_timer_start(hg_ctx *c)
{
    BIND(c);
    CU(c, cudaEventRecord(c->ev0, c->stream));
    return HG_OK;
}

int hg_timer_stop(hg_ctx *c, float *ms)
{
    BIND(c);
    NEED(c, ms, "elapsed_ms is NULL");
    CU(c, cudaEventRecord(c->ev1, c->stream));
    CU(c, cudaEventSynchronize(c->ev1));
    CU(c, cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return HG_OK;
}

int hg_ctx_set_sampling(hg_ctx *c, int sampling)
{
    if (!c) return HG_ERR_INVALID;
    NEED(c, sampling == HG_NEAREST || sampling == HG_BILINEAR, "sampling must be HG_NEAREST or HG_BILINEAR");
    c->sampling = sampling;
    return HG_OK;
}

int hg_launch_count(hg_ctx *c, uint64_t *n)
{
    if (!c || !n) return HG_ERR_INVALID;
    *n = c->launches;
    return HG_OK;
}

int hg_profile_enable(hg_ctx *c, int on)
{
    if (!c) return HG_ERR_INVALID;
    c->prof = on != 0;
    c->prof_used = 0;
    return HG_OK;
}

int hg_profile_read(hg_ctx *c, double *total_ms, uint64_t *n_kernels)
{
    BIND(c);
    NEED(c, total_ms && n_kernels, "NULL argument");
    CU(c, cudaStreamSynchronize(c->stream));
    double tot = 0.0;
    for (size_t i = 0; i < c->prof_used; ++i) {
        float ms = 0.f;
        CU(c, cudaEventElapsedTime(&ms, c->prof_ev[i].first, c->prof_ev[i].second));
        tot += ms;
    }
    *total_ms = tot;
    *n_kernels = c->prof_used;
    c->prof_used = 0;
    return HG_OK;
}

/* ------------------------------------------------------------------ image */
int hg_image_set(hg_ctx *c, const uint8_t *rgba, int w, int h)
{
    BIND(c);
    NEED(c, rgba, "rgba_host is NULL");
    TRY(check_image_dims(c, w, h));
    const size_t bytes = (size_t)w * h * 4;
    TRY(ensure(c, c->img_own, bytes));
    CU(c, cudaMemcpyAsync(c->img_own.p, rgba, bytes, cudaMemcpyHostToDevice, c->stream));
    c->img = (const uint32_t *)c->img_own.p;
    c->W = w;
    c->H = h;
    c->img_tm_ok = encode_tmaps(c, c->img, w, h, c->img_tm);
    return HG_OK;
}

int hg_image_set_device(hg_ctx *c, const void *rgba_dev, int w, int h)
{
    BIND(c);
    NEED(c, rgba_dev, "rgba_dev is NULL");
    NEED(c, ((uintptr_t)rgba_dev & 3) == 0, "rgba_dev must be 4-byte aligned");
    TRY(check_image_dims(c, w, h));
    c->img = (const uint32_t *)rgba_dev;
    c->W = w;
    c->H = h;
    c->img_tm_ok = encode_tmaps(c, c->img, w, h, c->img_tm);
    return HG_OK;
}

/* ------------------------------------------------------------------ solves */
int hg_solve_affine(hg_ctx *c, const double src[6], const double dst[6], float out[6])
{
    BIND(c);
    NEED(c, src && dst && out, "NULL argument");
    TRY(upload_small(c, SC_SRC, src, 48));
    TRY(upload_small(c, SC_DST, dst, 48));
    SolveArgs a{};
    a.src = (const double *)((char *)c->scratch.p + SC_SRC);
    a.dst = (const double *)((char *)c->scratch.p + SC_DST);
    a.out_f = (float *)((char *)c->scratch.p + SC_MAT);
    a.n = 1;
    a.op = 0;
    TRY(launch_solve(c, a));
    TRY(download_small(c, out, SC_MAT, 24));
    CU(c, cudaStreamSynchronize(c->stream));
    return HG_OK;
}

int hg_solve_projective(hg_ctx *c, const double src[8], const double dst[8], double out[8])
{
    BIND(c);
    NEED(c, src && dst && out, "NULL argument");
    TRY(upload_small(c, SC_SRC, src, 64));
    TRY(upload_small(c, SC_DST, dst, 64));
    SolveArgs a{};
    a.src = (const double *)((char *)c->scratch.p + SC_SRC);
    a.dst = (const double *)((char *)c->scratch.p + SC_DST);
    a.out_d = (double *)((char *)c->scratch.p + SC_MAT);
    a.n = 1;
    a.op = 1;
    TRY(launch_solve(c, a));
    TRY(download_small(c, out, SC_MAT, 64));
    CU(c, cudaStreamSynchronize(c->stream));
    return HG_OK;
}

int hg_inverse_affine(hg_ctx *c, const float m[6], float out[6])
{
    BIND(c);
    NEED(c, m && out, "NULL argument");
    TRY(upload_small(c, SC_MISC, m, 24));
    SolveArgs a{};
    a.in_f = (const float *)((char *)c->scratch.p + SC_MISC);
    a.out_f = (float *)((char *)c->scratch.p + SC_MAT);
    a.n = 1;
    a.op = 2;
    TRY(launch_solve(c, a));
    TRY(download_small(c, out, SC_MAT, 24));
    CU(c, cudaStreamSynchronize(c->stream));
    return HG_OK;
}

int hg_transform_limits(hg_ctx *c, int kind, const void *matrix, double w, double h, double out[4])
{
    BIND(c);
    NEED(c, matrix && out, "NULL argument");
    NEED(c, kind == HG_AFFINE || kind == HG_PROJECTIVE, "kind must be HG_AFFINE or HG_PROJECTIVE");
    TRY(upload_small(c, SC_MAT, matrix, kind == HG_AFFINE ? 24 : 64));
    limits_kernel<<<1, 32, 0, c->stream>>>(kind, (char *)c->scratch.p + SC_MAT, w, h,
                                           (double *)((char *)c->scratch.p + SC_LIM));
    c->launches++;
    CU(c, cudaGetLastError());
    TRY(download_small(c, out, SC_LIM, 32));
    CU(c, cudaStreamSynchronize(c->stream));
    return HG_OK;
}

int hg_solve_with_limits(hg_ctx *c, int kind, const double *src, const double *dst, double w, double h,
                         void *matrix_out, double limits_out[4])
{
    BIND(c);
    NEED(c, src && dst && matrix_out && limits_out, "NULL argument");
    NEED(c, kind == HG_AFFINE || kind == HG_PROJECTIVE, "kind must be HG_AFFINE or HG_PROJECTIVE");
    const size_t pb = kind == HG_AFFINE ? 48 : 64;
    const size_t mb = kind == HG_AFFINE ? 24 : 64;
    if (!c->stream_aux) {
        CU(c, cudaStreamCreateWithFlags(&c->stream_aux, cudaStreamNonBlocking));
        TRY(ensure(c, c->scratch_aux, 4096));
    }
    cudaStream_t st = c->stream_aux;
    char *sc = (char *)c->scratch_aux.p;
    CU(c, cudaMemcpyAsync(sc + SC_SRC, src, pb, cudaMemcpyHostToDevice, st));
    CU(c, cudaMemcpyAsync(sc + SC_DST, dst, pb, cudaMemcpyHostToDevice, st));
    SolveArgs a{};
    a.src = (const double *)(sc + SC_SRC);
    a.dst = (const double *)(sc + SC_DST);
    a.out_f = (float *)(sc + SC_MAT);
    a.out_d = (double *)(sc + SC_MAT);
    a.n = 1;
    a.op = kind == HG_AFFINE ? 0 : 1;
    TRY(launch_solve(c, a, st));
    limits_kernel<<<1, 32, 0, st>>>(kind, sc + SC_MAT, w, h, (double *)(sc + SC_LIM));
    c->launches++;
    CU(c, cudaGetLastError());
    // one D2H for matrix + limits (contiguous in scratch: [SC_MAT, SC_LIM+32)), into the upper half of the pinned page
    char *pin = (char *)c->pinned + 2048;
    CU(c, cudaMemcpyAsync(pin, sc + SC_MAT, SC_LIM + 32 - SC_MAT, cudaMemcpyDeviceToHost, st));
    CU(c, cudaStreamSynchronize(st));
    memcpy(matrix_out, pin, mb);
    memcpy(limits_out, pin + (SC_LIM - SC_MAT), 32);
    return HG_OK;
}

/* ------------------------------------------------------------------ affine / projective warps */
static int warp_inverse_common(hg_ctx *c, int kind, const void *inv_host, bool solve_on_device, int x_off,
                               int y_off, int o_w, int o_h, uint8_t *out_host, void *out_dev, bool tall_hint = false)
{
    NEED(c, kind == HG_AFFINE || kind == HG_PROJECTIVE, "kind must be HG_AFFINE or HG_PROJECTIVE");
    if (!c->img) return fail(c, HG_ERR_STATE, "no image set (hg_image_set)");
    TRY(check_window(c, x_off, y_off, o_w, o_h));
    const size_t bytes = (size_t)o_w * o_h * 4;
    uint32_t *dst = nullptr;
    TRY(pick_out(c, out_dev, bytes, &dst));
    GeoParams P{};
    P.one.src = c->img;
    P.one.out = dst;
    P.one.W = c->W;
    P.one.H = c->H;
    P.one.xOff = x_off;
    P.one.yOff = y_off;
    P.one.oW = o_w;
    P.one.oH = o_h;
    P.many = nullptr;
    if (solve_on_device) {
        P.mats_dev = (char *)c->scratch.p + SC_MAT;
    } else {
        P.mats_dev = nullptr;
        if (kind == HG_AFFINE)
            for (int k = 0; k < 6; ++k) P.mat_val[k] = (double)((const float *)inv_host)[k];
        else
            for (int k = 0; k < 8; ++k) P.mat_val[k] = ((const double *)inv_host)[k];
    }
    P.has_tm = c->img_tm_ok ? 1 : 0;
    if (c->img_tm_ok) memcpy(P.tm_val, c->img_tm, sizeof c->img_tm);
    const bool tall = solve_on_device ? tall_hint : map_is_rotated(kind, P.mat_val, x_off, y_off, o_w, o_h);
    TRY(launch_geo(c, kind, P, o_w, o_h, 1, c->img_tm_ok, c->stream, tall));
    return finish_out(c, dst, bytes, out_host);
}

int hg_warp_inverse_matrix(hg_ctx *c, int kind, const void *inv_matrix, int x_off, int y_off, int o_w, int o_h,
                           uint8_t *out_host, void *out_dev)
{
    BIND(c);
    NEED(c, inv_matrix, "inv_matrix is NULL");
    return warp_inverse_common(c, kind, inv_matrix, false, x_off, y_off, o_w, o_h, out_host, out_dev);
}

int hg_warp_inverse_points(hg_ctx *c, int kind, const double *dst_pts, const double *src_pts, int x_off,
                           int y_off, int o_w, int o_h, uint8_t *out_host, void *out_dev)
{
    BIND(c);
    NEED(c, dst_pts && src_pts, "NULL points");
    NEED(c, kind == HG_AFFINE || kind == HG_PROJECTIVE, "kind must be HG_AFFINE or HG_PROJECTIVE");
    const size_t pb = kind == HG_AFFINE ? 48 : 64;
    // inverse matrix = calculateTransformMatrix(kind, dstPoints, srcPoints)  (H.js:994).  The points travel as kernel
    // parameters: a copy from pageable memory would make the host wait for everything queued on the stream — the image
    // upload of a preceding hg_image_set — before the solve and the pixel loop could even be queued behind it.
    SolvePoints sp{};
    memcpy(sp.src, dst_pts, pb);
    memcpy(sp.dst, src_pts, pb);
    solve_points_kernel<<<1, 32, 0, c->stream>>>(sp, kind == HG_AFFINE ? 0 : 1, (float *)((char *)c->scratch.p + SC_MAT),
                                                 (double *)((char *)c->scratch.p + SC_MAT));
    c->launches++;
    CU(c, cudaGetLastError());
    // inverse map: dst -> src
    return warp_inverse_common(c, kind, nullptr, true, x_off, y_off, o_w, o_h, out_host, out_dev, points_map_is_rotated(dst_pts, src_pts));
}

// a winner plane of `ints` entries, all -1 (the gather pass resets what it reads, so a fill is only needed after the
// buffer grew or a failed call left it dirty)
static int ensure_winner(hg_ctx *c, size_t ints)
{
    if (sizeof(int) * ints <= c->winner.cap && c->winner_clean) return HG_OK;
    TRY(ensure(c, c->winner, sizeof(int) * ints));
    const long long n = (long long)(c->winner.cap / sizeof(int));
    fill_minus_one_kernel<<<(unsigned)(c->sm_count * 8), 256, 0, c->stream>>>((int *)c->winner.p, n);
    c->launches++;
    CU(c, cudaGetLastError());
    c->winner_clean = true;
    return HG_OK;
}

// Does the forward affine matrix map the pixel lattice onto itself one to one (forward.cuh, "lattice")?  Linear part a
// signed permutation with exact 0 / +-1 entries; translation e with e == 0 or 2^-8 <= |e| <= 2^18, so that
// (+-x) + e and the subtraction of the offset are exact in double and round((+-x) + e - xOff) = +-x + round(e - xOff);
// and no source pixel of the loop domain lands outside [0, oW) in x (the flat-index wrap of Q3 would otherwise fold
// several (nx, ny) onto one output pixel).  Fills the plan fields of `a` when it does.
static bool forward_lattice_plan(FwdArgs &a)
{
    a.lattice = 0;
    if (a.kind != HG_AFFINE) return false;
    const double m0 = a.mat[0], m1 = a.mat[1], m2 = a.mat[2], m3 = a.mat[3], e = a.mat[4], f = a.mat[5];
    auto unit = [](double v) { return v == 0.0 || v == 1.0 || v == -1.0; };
    if (!unit(m0) || !unit(m1) || !unit(m2) || !unit(m3)) return false;
    const int sxx = (int)m0, syx = (int)m1, sxy = (int)m2, syy = (int)m3;
    if (std::abs(sxx) + std::abs(sxy) != 1 || std::abs(syx) + std::abs(syy) != 1 || std::abs(sxx) + std::abs(syx) != 1) return false;
    auto exact = [](double v) { return v == 0.0 || (std::fabs(v) >= 0.00390625 && std::fabs(v) <= 262144.0); };
    if (!exact(e) || !exact(f)) return false;
    if (a.W > 65536 || a.H > 65536) return false;
    auto round_half_up = [](double v) { const double fl = std::floor(v); return (long long)fl + ((v - fl) >= 0.5 ? 1 : 0); };
    const long long rx = round_half_up(e - (double)a.xOff), ry = round_half_up(f - (double)a.yOff);
    // nx over the loop domain x in [0, W), y in [0, H)
    const long long xs[2] = {0, a.W - 1}, ys[2] = {0, a.H - 1};
    long long nmin = (1LL << 62), nmax = -(1LL << 62);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
            const long long nx = sxx * xs[i] + sxy * ys[j] + rx;
            if (nx < nmin) nmin = nx;
            if (nx > nmax) nmax = nx;
        }
    if (nmin < 0 || nmax >= a.oW) return false;
    if (std::llabs(rx) > (1LL << 20) || std::llabs(ry) > (1LL << 20)) return false;
    a.lattice = 1;
    a.ixx = sxx; a.ixy = syx; a.iyx = sxy; a.iyy = syy;  // inverse of a signed permutation = its transpose
    a.rx = (int)rx; a.ry = (int)ry;
    a.shift_copy = (sxx == 1 && syy == 1 && (a.oW & 3) == 0 && (a.W & 3) == 0 && ((uintptr_t)a.src & 15) == 0) ? 1 : 0;
    return true;
}

static unsigned fwd_blocks(hg_ctx *c, long long items, int n_frames)
{
    long long blocks = (items + 255) / 256;
    long long cap = (long long)c->sm_count * (n_frames > 1 ? 4 : 16);
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks < 1 ? 1 : blocks);
}

static int run_forward(hg_ctx *c, FwdArgs &a, bool piecewise, uint32_t *dst, size_t bytes, uint8_t *out_host)
{
    const long long npix = (long long)a.oW * a.oH;
    a.out = dst;
    FwdParams P{};
    P.many = nullptr;
    if (!piecewise && forward_lattice_plan(a)) {
        P.one = a;
        TRY(prof_begin(c));
        forward_lattice_kernel<<<dim3(fwd_blocks(c, (npix + 3) / 4, 1), 1), 256, 0, c->stream>>>(P);
        c->launches++;
        CU(c, cudaGetLastError());
        TRY(prof_end(c));
        return finish_out(c, dst, bytes, out_host);
    }
    TRY(ensure_winner(c, (size_t)npix));
    a.winner = (int *)c->winner.p;
    P.one = a;
    c->winner_clean = false;
    const long long n = (long long)a.domW * a.domH;
    TRY(prof_begin(c));
    if (n > 0) {
        const dim3 g(fwd_blocks(c, n, 1), 1);
        if (piecewise) forward_scatter_kernel<2><<<g, 256, 0, c->stream>>>(P);
        else if (a.kind == 0) forward_scatter_kernel<0><<<g, 256, 0, c->stream>>>(P);
        else forward_scatter_kernel<1><<<g, 256, 0, c->stream>>>(P);
        c->launches++;
        CU(c, cudaGetLastError());
    }
    forward_gather_kernel<<<dim3(fwd_blocks(c, (npix + 3) / 4, 1), 1), 256, 0, c->stream>>>(P);
    c->launches++;
    CU(c, cudaGetLastError());
    TRY(prof_end(c));
    c->winner_clean = true;
    return finish_out(c, dst, bytes, out_host);
}

// a batch of forward frames whose FwdArgs (winner still unset) are in `fa`: lattice frames in one launch, the others
// through a small ring of winner planes that stays resident in L2
static int run_forward_batch(hg_ctx *c, std::vector<FwdArgs> &fa, bool piecewise)
{
    const int n_frames = (int)fa.size();
    long long max_npix = 1, max_dom = 1;
    int n_lattice = 0;
    for (auto &a : fa) {
        if (!piecewise && forward_lattice_plan(a)) { n_lattice++; continue; }
        const long long npix = (long long)a.oW * a.oH, dom = (long long)a.domW * a.domH;
        if (npix > max_npix) max_npix = npix;
        if (dom > max_dom) max_dom = dom;
    }
    max_npix = (max_npix + 3) & ~3LL;  // planes start 16-byte aligned (the gather pass reads and resets them four at a time)
    // planes: as many as fit ~48 MB (they are read, reset and re-used while L2-resident), between 1 and 8
    int planes = (int)(((long long)c->fwd_plane_mb << 20) / (max_npix * 4));
    if (planes < 1) planes = 1;
    if (planes > 8) planes = 8;
    // two lanes of sub-batches on two streams, each with its own half of the planes: the scatter of one sub-batch (bound by
    // the atomic rate) runs beside the gather of the other (bound by its dependent loads)
    const bool two_lanes = planes >= 2;
    const int half = two_lanes ? planes / 2 : planes;
    if (n_lattice < n_frames) {
        TRY(ensure_winner(c, (size_t)max_npix * planes));
        if (two_lanes && !c->stream2) {
            CU(c, cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
            CU(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
            CU(c, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
        }
        // sub-batch i = general frames [i * half, (i + 1) * half) in frame order; it uses planes (i % 2) * half + j
        int k = 0;
        for (auto &a : fa)
            if (!a.lattice) {
                const int sub = k / half, j = k % half;
                a.winner = (int *)c->winner.p + (size_t)max_npix * ((two_lanes ? (sub & 1) * half : 0) + j);
                ++k;
            }
    }
    TRY(ensure(c, c->fwd_args, sizeof(FwdArgs) * (size_t)n_frames));
    // general frames first in device order? no: keep frame order, kernels skip frames of the other kind
    CU(c, cudaMemcpyAsync(c->fwd_args.p, fa.data(), sizeof(FwdArgs) * (size_t)n_frames, cudaMemcpyHostToDevice, c->stream));
    TRY(prof_begin(c));
    if (n_lattice > 0) {
        long long max_q = 1;
        for (auto &a : fa)
            if (a.lattice && ((long long)a.oW * a.oH + 3) / 4 > max_q) max_q = ((long long)a.oW * a.oH + 3) / 4;
        for (int f0 = 0; f0 < n_frames; f0 += 32768) {
            const int nf = n_frames - f0 < 32768 ? n_frames - f0 : 32768;
            FwdParams P{};
            P.many = (const FwdArgs *)c->fwd_args.p + f0;
            forward_lattice_kernel<<<dim3(fwd_blocks(c, max_q, nf), (unsigned)nf), 256, 0, c->stream>>>(P);
            c->launches++;
        }
        CU(c, cudaGetLastError());
    }
    if (n_lattice < n_frames) {
        c->winner_clean = false;
        if (two_lanes) {
            CU(c, cudaEventRecord(c->ev_fork, c->stream));   // frame descriptors uploaded, planes clean
            CU(c, cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
        }
        // sub-batches of consecutive frames that together hold exactly `half` general frames (the last one fewer)
        int f0 = 0, sub = 0;
        while (f0 < n_frames) {
            int f1 = f0, g = 0;
            while (f1 < n_frames && f1 - f0 < 32768 && (g < half || fa[(size_t)f1].lattice)) {
                if (!fa[(size_t)f1].lattice) g++;
                f1++;
            }
            if (g > 0) {
                const int nf = f1 - f0;
                cudaStream_t st = (two_lanes && (sub & 1)) ? c->stream2 : c->stream;
                FwdParams P{};
                P.many = (const FwdArgs *)c->fwd_args.p + f0;
                const dim3 gs(fwd_blocks(c, max_dom, nf), (unsigned)nf), gg(fwd_blocks(c, (max_npix + 3) / 4, nf), (unsigned)nf);
                if (piecewise) forward_scatter_kernel<2><<<gs, 256, 0, st>>>(P);
                else if (fa[0].kind == 0) forward_scatter_kernel<0><<<gs, 256, 0, st>>>(P);   // one kind per batch
                else forward_scatter_kernel<1><<<gs, 256, 0, st>>>(P);
                forward_gather_kernel<<<gg, 256, 0, st>>>(P);
                c->launches += 2;
                ++sub;
            }
            f0 = f1;
        }
        CU(c, cudaGetLastError());
        if (two_lanes) {
            CU(c, cudaEventRecord(c->ev_join, c->stream2));
            CU(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
        }
        c->winner_clean = true;
    }
    TRY(prof_end(c));
    return HG_OK;
}

int hg_warp_forward_matrix(hg_ctx *c, int kind, const void *fwd_matrix, int x_off, int y_off, int o_w, int o_h,
                           uint8_t *out_host, void *out_dev)
{
    BIND(c);
    NEED(c, fwd_matrix, "fwd_matrix is NULL");
    NEED(c, kind == HG_AFFINE || kind == HG_PROJECTIVE, "kind must be HG_AFFINE or HG_PROJECTIVE");
    if (!c->img) return fail(c, HG_ERR_STATE, "no image set (hg_image_set)");
    TRY(check_window(c, x_off, y_off, o_w, o_h));
    const size_t bytes = (size_t)o_w * o_h * 4;
    uint32_t *dst = nullptr;
    TRY(pick_out(c, out_dev, bytes, &dst));
    FwdArgs a{};
    a.src = c->img;
    a.kind = kind;
    if (kind == HG_AFFINE)
        for (int k = 0; k < 6; ++k) a.mat[k] = (double)((const float *)fwd_matrix)[k];
    else
        for (int k = 0; k < 8; ++k) a.mat[k] = ((const double *)fwd_matrix)[k];
    a.W = c->W; a.H = c->H;
    a.xOff = x_off; a.yOff = y_off; a.oW = o_w; a.oH = o_h;
    a.minX = 0; a.minY = 0; a.domW = c->W; a.domH = c->H;  // for (y < H) for (x < W), H.js:919-920
    return run_forward(c, a, false, dst, bytes, out_host);
}

int hg_warp_inverse_batch(hg_ctx *c, int kind, const void *inv_matrices, const hg_frame *frames, int n_frames)
{
    BIND(c);
    NEED(c, inv_matrices && frames, "NULL argument");
    NEED(c, kind == HG_AFFINE || kind == HG_PROJECTIVE, "kind must be HG_AFFINE or HG_PROJECTIVE");
    NEED(c, n_frames >= 1, "n_frames must be >= 1");
    std::vector<GeoFrame> gf((size_t)n_frames);
    int max_ow = 1, max_oh = 1;
    bool staged = false;
    for (int f = 0; f < n_frames; ++f) {
        const hg_frame &h = frames[f];
        TRY(check_window(c, h.x_off, h.y_off, h.o_w, h.o_h));
        NEED(c, h.out_dev && ((uintptr_t)h.out_dev & 15) == 0, "frame out_dev must be a 16-byte aligned device pointer");
        GeoFrame &g = gf[(size_t)f];
        if (h.src_dev) {
            TRY(check_image_dims(c, h.src_w, h.src_h));
            g.src = (const uint32_t *)h.src_dev;
            g.W = h.src_w;
            g.H = h.src_h;
        } else {
            if (!c->img) return fail(c, HG_ERR_STATE, "no image set (hg_image_set)");
            g.src = c->img;
            g.W = c->W;
            g.H = c->H;
        }
        TRY(tmaps_device(c, g.src, g.W, g.H, &g.tm));
        staged = staged || g.tm != nullptr;
        g.out = (uint32_t *)h.out_dev;
        g.xOff = h.x_off;
        g.yOff = h.y_off;
        g.oW = h.o_w;
        g.oH = h.o_h;
        if (h.o_w > max_ow) max_ow = h.o_w;
        if (h.o_h > max_oh) max_oh = h.o_h;
    }
    const size_t mstride = kind == HG_AFFINE ? 24 : 64;
    TRY(ensure(c, c->frames, sizeof(GeoFrame) * (size_t)n_frames));
    TRY(ensure(c, c->mats, mstride * (size_t)n_frames));
    CU(c, cudaMemcpyAsync(c->frames.p, gf.data(), sizeof(GeoFrame) * (size_t)n_frames, cudaMemcpyHostToDevice, c->stream));
    // pageable sources: cudaMemcpyAsync returns once they are staged, so gf may die at scope exit
    CU(c, cudaMemcpyAsync(c->mats.p, inv_matrices, mstride * (size_t)n_frames, cudaMemcpyHostToDevice, c->stream));
    // blockIdx.y of the direct kernel and the 32-bit tile index of the staged kernel bound the frames per launch
    int chunk = 32768;
    if (staged) {
        const long long tpf = (long long)geo_tiles_x(max_ow) * geo_tiles_y(max_oh, 1);
        const long long fit = (1LL << 30) / tpf;
        if (fit < chunk) chunk = fit < 1 ? 1 : (int)fit;
    }
    for (int f0 = 0; f0 < n_frames; f0 += chunk) {
        const int nf = n_frames - f0 < chunk ? n_frames - f0 : chunk;
        GeoParams P{};
        P.many = (const GeoFrame *)c->frames.p + f0;
        P.mats_dev = (const char *)c->mats.p + mstride * (size_t)f0;
        double m0[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (kind == HG_AFFINE)
            for (int k = 0; k < 6; ++k) m0[k] = (double)((const float *)inv_matrices)[6 * (size_t)f0 + k];
        else
            for (int k = 0; k < 8; ++k) m0[k] = ((const double *)inv_matrices)[8 * (size_t)f0 + k];
        const bool tall = map_is_rotated(kind, m0, frames[f0].x_off, frames[f0].y_off, frames[f0].o_w, frames[f0].o_h);
        TRY(launch_geo(c, kind, P, max_ow, max_oh, nf, staged, c->stream, tall));
    }
    return HG_OK;
}

/* ------------------------------------------------------------------ triangulation */
int hg_delaunay(const double *points, int n_points, uint32_t *triangles_out, int capacity_triangles, int *n_triangles)
{
    if (!points || !n_triangles || n_points < 0 || capacity_triangles < 0 || (capacity_triangles > 0 && !triangles_out))
        return HG_ERR_INVALID;
    *n_triangles = 0;
    try {  // nothing may unwind through the C ABI
        const std::vector<uint32_t> t = hg_delaunay_detail::triangulate(points, (size_t)n_points);
        const size_t nt = t.size() / 3;
        if (nt > (size_t)capacity_triangles) return HG_ERR_INVALID;
        if (nt) memcpy(triangles_out, t.data(), t.size() * sizeof(uint32_t));
        *n_triangles = (int)nt;
        return HG_OK;
    } catch (const std::bad_alloc &) {
        return HG_ERR_NOMEM;
    } catch (...) {
        return HG_ERR_INVALID;
    }
}

/* ------------------------------------------------------------------ image files */
int hg_png_decode(const uint8_t *png, size_t png_bytes, uint8_t *rgba_out, size_t capacity_bytes, int *w, int *h)
{
    if (!png || !w || !h) return HG_ERR_INVALID;
    *w = *h = 0;
    try {  // the bytes are untrusted: nothing may unwind through the C ABI
        hg_png_detail::Header hd;
        if (hg_png_detail::decode(png, png_bytes, hd, nullptr)) return HG_ERR_INVALID;  // header + chunk CRCs
        *w = (int)hd.w;
        *h = (int)hd.h;
        if (!rgba_out) return HG_OK;
        if (capacity_bytes < (size_t)hd.w * hd.h * 4) return HG_ERR_INVALID;
        return hg_png_detail::decode(png, png_bytes, hd, rgba_out) ? HG_ERR_INVALID : HG_OK;
    } catch (const std::bad_alloc &) {
        return HG_ERR_NOMEM;
    } catch (...) {
        return HG_ERR_INVALID;
    }
}

int hg_jpeg_decode(const uint8_t *jpg, size_t jpg_bytes, uint8_t *rgba_out, size_t capacity_bytes, int *w, int *h)
{
    if (!jpg || !w || !h) return HG_ERR_INVALID;
    *w = *h = 0;
    try {  // the bytes are untrusted: nothing may unwind through the C ABI
        int ww = 0, hh = 0;
        int r = hg_jpeg_detail::decode(jpg, jpg_bytes, ww, hh, nullptr);  // frame header
        if (r) return r == hg_jpeg_detail::UNSUPPORTED ? HG_ERR_UNSUPPORTED : HG_ERR_INVALID;
        *w = ww;
        *h = hh;
        if (!rgba_out) return HG_OK;
        if (capacity_bytes < (size_t)ww * hh * 4) return HG_ERR_INVALID;
        int w2 = 0, h2 = 0;
        r = hg_jpeg_detail::decode(jpg, jpg_bytes, w2, h2, rgba_out);
        if (r == 0 && (w2 != ww || h2 != hh)) return HG_ERR_INVALID;  // the two passes must agree on the picture's size
        return r == 0 ? HG_OK : (r == hg_jpeg_detail::UNSUPPORTED ? HG_ERR_UNSUPPORTED : HG_ERR_INVALID);
    } catch (const std::bad_alloc &) {
        return HG_ERR_NOMEM;
    } catch (...) {
        return HG_ERR_INVALID;
    }
}

size_t hg_png_encode_bound(int w, int h)
{
    if (w < 1 || h < 1) return 0;
    const size_t z = (size_t)compressBound((uLong)(((size_t)w * 4 + 1) * (size_t)h));
    return 8 + 25 + 12 + z + 12 * (z >> 30) + 12;  // signature, IHDR, IDAT chunk(s) of <= 2^30 bytes, IEND
}

int hg_png_encode(const uint8_t *rgba, int w, int h, uint8_t *png_out, size_t capacity_bytes, size_t *png_bytes)
{
    if (!rgba || !png_out || !png_bytes || w < 1 || h < 1 || w > 65536 || h > 65536) return HG_ERR_INVALID;
    try {
        std::vector<uint8_t> out;
        if (hg_png_detail::encode(rgba, (uint32_t)w, (uint32_t)h, out)) return HG_ERR_INVALID;
        if (out.size() > capacity_bytes) return HG_ERR_INVALID;
        memcpy(png_out, out.data(), out.size());
        *png_bytes = out.size();
        return HG_OK;
    } catch (const std::bad_alloc &) {
        return HG_ERR_NOMEM;
    } catch (...) {
        return HG_ERR_INVALID;
    }
}

/* ------------------------------------------------------------------ piecewise */
static int check_points(hg_ctx *c, const float *p, int n, const char *what)
{
    for (int i = 0; i < 2 * n; ++i) {
        const float v = p[i];
        if (!(v >= -1048576.f && v <= 1048576.f))  // also rejects NaN / Inf
            return fail(c, HG_ERR_UNSUPPORTED, "%s[%d] = %g: piecewise points must be finite and |v| <= 2^20", what, i, (double)v);
    }
    return HG_OK;
}

static int check_point_floats(hg_ctx *c, const float *p, size_t n_floats, const char *what)
{
    // blocks of branch-free compares (the compiler vectorises them): a batch carries millions of coordinates and this
    // runs on the host in front of every launch; only a failing block is searched for the offender
    for (size_t b0 = 0; b0 < n_floats; b0 += 4096) {
        const size_t b1 = b0 + 4096 < n_floats ? b0 + 4096 : n_floats;
        unsigned bad = 0;
        for (size_t i = b0; i < b1; ++i) bad |= (unsigned)!(p[i] >= -1048576.f && p[i] <= 1048576.f);  // also NaN / Inf
        if (bad)
            for (size_t i = b0; i < b1; ++i)
                if (!(p[i] >= -1048576.f && p[i] <= 1048576.f))
                    return fail(c, HG_ERR_UNSUPPORTED, "%s[%zu] = %g: piecewise points must be finite and |v| <= 2^20", what, i,
                                (double)p[i]);
    }
    return HG_OK;
}

int hg_piecewise_set_mesh(hg_ctx *c, const float *src_pts, int n_pts, const uint32_t *tris, int n_tris)
{
    BIND(c);
    NEED(c, src_pts && tris, "NULL argument");
    NEED(c, n_pts >= 3 && n_tris >= 0, "need >= 3 points");
    for (int i = 0; i < 3 * n_tris; ++i)
        if (tris[i] >= (uint32_t)n_pts) return fail(c, HG_ERR_INVALID, "triangle index %u out of range", tris[i]);
    TRY(check_points(c, src_pts, n_pts, "src_pts"));
    TRY(ensure(c, c->src_pts, sizeof(float) * 2 * (size_t)n_pts));
    TRY(ensure(c, c->tris, sizeof(uint32_t) * 3 * (size_t)(n_tris > 0 ? n_tris : 1)));
    TRY(ensure(c, c->rec, sizeof(TriRec) * (size_t)(n_tris > 0 ? n_tris : 1)));
    CU(c, cudaMemcpyAsync(c->src_pts.p, src_pts, sizeof(float) * 2 * (size_t)n_pts, cudaMemcpyHostToDevice, c->stream));
    if (n_tris)
        CU(c, cudaMemcpyAsync(c->tris.p, tris, sizeof(uint32_t) * 3 * (size_t)n_tris, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    c->n_pts = n_pts;
    c->n_tris = n_tris;
    return HG_OK;
}

int hg_piecewise_mesh_size(hg_ctx *c, int *n_pts, int *n_tris)
{
    if (!c || !n_pts || !n_tris) return HG_ERR_INVALID;
    *n_pts = c->n_pts;
    *n_tris = c->n_tris;
    return HG_OK;
}

static int upload_dst_points(hg_ctx *c, const float *dst_pts, size_t n_frames)
{
    const size_t bytes = sizeof(float) * 2 * (size_t)c->n_pts * n_frames;
    TRY(ensure(c, c->dst_pts, bytes));
    CU(c, cudaMemcpyAsync(c->dst_pts.p, dst_pts, bytes, cudaMemcpyHostToDevice, c->stream));
    return HG_OK;
}

static int launch_setup(hg_ctx *c, const float *dst_dev, const float *map_pts_dev, float *fwd_out, float *inv_out)
{
    if (c->n_tris == 0) return HG_OK;
    PwSetupArgs a{};
    a.src_pts = (const float *)c->src_pts.p;
    a.dst_pts = dst_dev;
    a.map_pts = map_pts_dev;
    a.tris = (const uint32_t *)c->tris.p;
    a.rec = (TriRec *)c->rec.p;
    a.fwd_out = fwd_out;
    a.inv_out = inv_out;
    a.n_tris = c->n_tris;
    a.dst_stride = 0;
    a.rec_stride = 0;
    pw_setup_kernel<<<dim3((c->n_tris + 127) / 128, 1), 128, 0, c->stream>>>(a);
    c->launches++;
    CU(c, cudaGetLastError());
    return HG_OK;
}

static int launch_fill(hg_ctx *c, double map_width, double y_offset, long long map_len)
{
    TRY(ensure(c, c->map32, sizeof(int) * (size_t)(map_len > 0 ? map_len : 1)));
    if (map_len > 0) CU(c, cudaMemsetAsync(c->map32.p, 0xFF, sizeof(int) * (size_t)map_len, c->stream));
    c->map_len = map_len;
    if (c->n_tris == 0 || map_len <= 0) return HG_OK;
    PwFillArgs a{};
    a.rec = (const TriRec *)c->rec.p;
    a.map32 = (int *)c->map32.p;
    a.map_len = map_len;
    a.map_width = map_width;
    a.y_offset = y_offset;
    a.n_tris = c->n_tris;
    // rows of one triangle are spread over row_split blocks; any value is correct
    double rows_guess = map_width > 0 ? (double)map_len / map_width : 1.0;
    int split = (int)(rows_guess / 64.0) + 1;
    if (split > 64) split = 64;
    if ((long long)split * c->n_tris > (1 << 20)) split = (1 << 20) / c->n_tris + 1;
    a.row_split = split;
    pw_fill_kernel<<<dim3((unsigned)c->n_tris, (unsigned)split), dim3(32, 8), 0, c->stream>>>(a);
    c->launches++;
    CU(c, cudaGetLastError());
    return HG_OK;
}

int hg_piecewise_matrices(hg_ctx *c, const float *dst_pts, float *fwd_out, float *inv_out)
{
    BIND(c);
    NEED(c, dst_pts, "dst_pts is NULL");
    if (c->n_pts == 0) return fail(c, HG_ERR_STATE, "no mesh set (hg_piecewise_set_mesh)");
    TRY(check_points(c, dst_pts, c->n_pts, "dst_pts"));
    TRY(upload_dst_points(c, dst_pts, 1));
    const size_t mb = sizeof(float) * 6 * (size_t)(c->n_tris > 0 ? c->n_tris : 1);
    TRY(ensure(c, c->mats, 2 * mb));
    float *fd = (float *)c->mats.p, *id = (float *)((char *)c->mats.p + mb);
    TRY(launch_setup(c, (const float *)c->dst_pts.p, (const float *)c->dst_pts.p, fd, id));
    if (c->n_tris) {
        if (fwd_out) CU(c, cudaMemcpyAsync(fwd_out, fd, sizeof(float) * 6 * (size_t)c->n_tris, cudaMemcpyDeviceToHost, c->stream));
        if (inv_out) CU(c, cudaMemcpyAsync(inv_out, id, sizeof(float) * 6 * (size_t)c->n_tris, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(c, cudaStreamSynchronize(c->stream));
    return HG_OK;
}

int hg_piecewise_extents(hg_ctx *c, const float *dst_pts, int n_pts, int n_frames, double *out)
{
    BIND(c);
    NEED(c, dst_pts && out, "NULL argument");
    NEED(c, n_pts >= 1 && n_frames >= 1, "n_pts and n_frames must be >= 1");
    const size_t in_bytes = sizeof(float) * 2 * (size_t)n_pts * n_frames, out_bytes = sizeof(double) * 4 * (size_t)n_frames;
    TRY(ensure(c, c->dst_pts, in_bytes));
    TRY(ensure(c, c->mats, out_bytes));
    CU(c, cudaMemcpyAsync(c->dst_pts.p, dst_pts, in_bytes, cudaMemcpyHostToDevice, c->stream));
    pw_extent_kernel<<<(unsigned)((n_frames + 3) / 4), 128, 0, c->stream>>>((const float *)c->dst_pts.p, n_pts, n_frames,
                                                                          (double *)c->mats.p);
    c->launches++;
    CU(c, cudaGetLastError());
    CU(c, cudaMemcpyAsync(out, c->mats.p, out_bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return HG_OK;
}

int hg_build_index_map(hg_ctx *c, const float *pts, double map_width, double y_offset, int64_t map_len,
                       int16_t *map_out_host)
{
    BIND(c);
    NEED(c, pts, "pts is NULL");
    if (c->n_pts == 0) return fail(c, HG_ERR_STATE, "no mesh set (hg_piecewise_set_mesh)");
    NEED(c, map_len >= 0 && map_len < (1LL << 31), "map_len out of range");
    TRY(check_points(c, pts, c->n_pts, "pts"));
    TRY(upload_dst_points(c, pts, 1));
    TRY(launch_setup(c, nullptr, (const float *)c->dst_pts.p, nullptr, nullptr));
    TRY(launch_fill(c, map_width, y_offset, map_len));
    c->last_inv_pts.assign(pts, pts + 2 * (size_t)c->n_pts);
    c->last_inv_mw = map_width;
    c->last_inv_yoff = y_offset;
    c->last_inv_len = map_len;
    c->map32_current = true;
    if (map_out_host && map_len > 0) {
        TRY(ensure(c, c->map16, sizeof(short) * (size_t)map_len));
        map32_to_int16_kernel<<<(unsigned)((map_len + 255) / 256), 256, 0, c->stream>>>((const int *)c->map32.p,
                                                                                      (short *)c->map16.p, map_len);
        c->launches++;
        CU(c, cudaGetLastError());
        CU(c, cudaMemcpyAsync(map_out_host, c->map16.p, sizeof(short) * (size_t)map_len, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(c, cudaStreamSynchronize(c->stream));
    return HG_OK;
}

struct PwFrameHost {
    const uint32_t *src;
    uint32_t *out;
    int W, H, xOff, yOff, oW, oH;
};

// general (map-based) inverse piecewise warp of ONE frame whose destiny points are already on the device
static int pw_inverse_general_frame(hg_ctx *c, const float *dst_dev, const PwFrameHost &f, int min_src_x, int min_src_y)
{
    TRY(launch_setup(c, dst_dev, dst_dev, nullptr, nullptr));
    const long long map_len = (long long)f.oW * f.oH;
    TRY(launch_fill(c, (double)f.oW, (double)f.yOff, map_len));
    PwWarpArgs a{};
    a.src = f.src;
    a.out = f.out;
    a.map32 = (const int *)c->map32.p;
    a.rec = (const TriRec *)c->rec.p;
    a.W = f.W; a.H = f.H;
    a.xOff = f.xOff; a.yOff = f.yOff; a.oW = f.oW; a.oH = f.oH;
    a.minSrcX = min_src_x; a.minSrcY = min_src_y;
    a.n_tris = c->n_tris;
    TRY(prof_begin(c));
    pw_warp_inverse_kernel<<<grid_for(c, map_len, 1), 256, 0, c->stream>>>(a);
    c->launches++;
    CU(c, cudaGetLastError());
    TRY(prof_end(c));
    return HG_OK;
}

// per-chunk scratch of the fused path; lane 1 exists so that the binning passes of one chunk run beside the pixel kernel of
// the chunk before it
struct PwScratch {
    DevBuf *rec, *invd, *bin_cnt, *bin_ent, *bin_run, *fstatus, *fframes, *yr;
};
static PwScratch pw_scratch(hg_ctx *c, int lane)
{
    if (lane == 0) return PwScratch{&c->rec, &c->invd, &c->bin_cnt, &c->bin_ent, &c->bin_run, &c->fstatus, &c->fframes, &c->yr};
    return PwScratch{&c->rec_b, &c->invd_b, &c->bin_cnt_b, &c->bin_ent_b, &c->bin_run_b, &c->fstatus_b, &c->fframes_b, &c->yr_b};
}
// One band pass (pw_band_bins_kernel) or the span + run passes?  Measured on B200, 64 4K frames per call, whole step against
// the HBM bound: 162 triangles 0.475 (band) / 0.440 (span + runs); 7,938 triangles 0.212 / 0.242 — every band scans all row
// ranges, evaluates a row of a triangle once per band it can reach and runs at 24 warps per SM, which a fine mesh (two
// rows per lane and segment) does not repay.
static bool pw_band_binning(const hg_ctx *c)
{
    if (c->pw_binning) return c->pw_binning == 2;
    return c->n_tris <= 2048;
}
// Two lanes (the binning passes of chunk k+1 on a high-priority stream beside the pixel kernel of chunk k)?  Measured on B200:
// the band pass is bound by its instruction count and takes from the pixel kernel what it gains (config 3: 0.438 -> 0.413-0.429
// of the HBM bound), the span pass is bound by its returning atomics and shares the machine (config 4, 512 frames streamed in
// chunks of 64: 0.292 -> 0.301; chunks of 32: 0.298).
static bool pw_two_lanes(const hg_ctx *c) { return c->pw_lanes >= 0 ? c->pw_lanes == 1 : !pw_band_binning(c); }
static int pw_scratch_ensure(hg_ctx *c, const PwScratch &S, size_t T, int nF, size_t total_bins)
{
    TRY(ensure(c, *S.yr, sizeof(int2) * (T ? T : 1) * (size_t)nF));
    // + one row group of slack: the pixel kernel's bin pointers may step (and read, but never use) past the last row
    const size_t slack = (size_t)PWF_GROUP_ROWS * 1024;
    TRY(ensure(c, *S.rec, sizeof(TriRec) * (T ? T : 1) * (size_t)nF));
    TRY(ensure(c, *S.invd, sizeof(float) * 8 * (T ? T : 1) * (size_t)nF));
    TRY(ensure(c, *S.bin_cnt, sizeof(unsigned) * 2 * (total_bins + slack)));
    TRY(ensure(c, *S.bin_ent, sizeof(unsigned) * (total_bins + slack) * PW_BIN_CAP));
    TRY(ensure(c, *S.bin_run, sizeof(uint4) * 2 * (total_bins + slack)));
    TRY(ensure(c, *S.fstatus, sizeof(int) * (size_t)nF));
    TRY(ensure(c, *S.fframes, sizeof(FusedFrame) * (size_t)nF));
    return HG_OK;
}
// the two lanes of a pipelined piecewise call: binning passes on a high-priority stream of their own (their CTAs are
// placed as pixel CTAs of the previous chunk retire: the passes are bound by the latency of their atomics, the pixel
// kernel by L1 and HBM, so they share the machine well), pixel kernels on the context stream
static int pw_lanes_begin(hg_ctx *c)
{
    if (!c->stream_pre) {
        int lo = 0, hi = 0;
        CU(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(c, cudaStreamCreateWithPriority(&c->stream_pre, cudaStreamNonBlocking, hi));
        for (auto &e : c->ev_pre) CU(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto &e : c->ev_pix) CU(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    // everything queued on the context stream so far (image upload, mesh) is visible to the binning lane
    CU(c, cudaEventRecord(c->ev_pre[0], c->stream));
    CU(c, cudaStreamWaitEvent(c->stream_pre, c->ev_pre[0], 0));
    return HG_OK;
}

// the binning passes of the fused path over the nF FusedFrame descriptors in S.fframes (written by the host or by
// pw_stream_frames_kernel) on stream `st`; bin counters and status flags are already zeroed.  Grids are sized for a
// max_ow x max_oh window: CTAs beyond a smaller frame's extent exit at once.
static int pw_fused_prepass(hg_ctx *c, const PwScratch &S, cudaStream_t st, const float *dst_dev, int nF, int max_ow, int max_oh,
                            size_t max_bins, bool records_inline)
{
    const size_t T = (size_t)c->n_tris;
    const FusedFrame *frames = (const FusedFrame *)S.fframes->p;
    if (T > 0) {
        PwSetupArgs a{};
        a.src_pts = (const float *)c->src_pts.p;
        a.dst_pts = dst_dev;
        a.map_pts = dst_dev;
        a.tris = (const uint32_t *)c->tris.p;
        a.rec = (TriRec *)S.rec->p;
        a.invd_out = (float *)S.invd->p;
        a.yr_out = (int2 *)S.yr->p;
        a.n_tris = c->n_tris;
        a.dst_stride = 2 * (size_t)c->n_pts;
        a.rec_stride = T;
        pw_setup_kernel<<<dim3((unsigned)((T + 127) / 128), (unsigned)nF), 128, 0, st>>>(a);
        c->launches++;
        CU(c, cudaGetLastError());
    }
    if (pw_band_binning(c)) {
        // bins, overlaps and run records of a band of map rows in one pass through shared memory
        const int band_rows = pwb_band_rows(pwf_bins_x(max_ow));
        pw_band_bins_kernel<<<dim3((unsigned)((max_oh + band_rows - 1) / band_rows), (unsigned)nF), PWB_THREADS, PWB_SMEM, st>>>(frames);
        c->launches++;
        CU(c, cudaGetLastError());
        return HG_OK;
    }
    if (T > 0) {
        // lanes per triangle: a mesh of T triangles over oH rows has triangles ~ oH / sqrt(T/2) rows tall; about three rows
        // per lane keep the lanes busy without leaving the machine empty for coarse meshes
        const double rows_guess = (double)max_oh / sqrt((double)T / 2.0 + 1.0);
        int lpt_log2 = 3;
        while (lpt_log2 < 9 && (double)(1 << lpt_log2) * 3.0 < rows_guess) ++lpt_log2;
        // ... and a coarse mesh spreads its triangles further until the launch fills the machine about twice
        while (lpt_log2 < 9 && (double)(T << lpt_log2) * nF < 2.0 * 2048.0 * c->sm_count && (double)(1 << lpt_log2) < rows_guess) ++lpt_log2;
        const size_t span_threads = T << lpt_log2;
        const dim3 sg((unsigned)((span_threads + 127) / 128), (unsigned)nF);
        if (rows_guess >= 96.0) pw_span_bin_kernel<true><<<sg, 128, 0, st>>>(frames, lpt_log2);
        else pw_span_bin_kernel<false><<<sg, 128, 0, st>>>(frames, lpt_log2);
        c->launches++;
        CU(c, cudaGetLastError());
    }
    // the run records: a pass of their own, unless every frame of the launch has a width that is a multiple of four — then
    // the pixel kernel builds the records of its rows in shared memory itself (no record array written and read back)
    // (measured on B200, 64 4K frames per launch: building them inside the pixel kernel costs the 162-triangle mesh as much
    // as the pass it saves and the 7,938-triangle mesh 11 % more — the record builder's instructions land in a kernel that is
    // already bound by its instruction and L1 rates — so it is opt-in: HG_PWF_RECORDS_INLINE=1)
    if (!records_inline) {
        pw_bin_runs_kernel<<<dim3((unsigned)((max_bins + 127) / 128), (unsigned)nF), 128, 0, st>>>(frames);
        c->launches++;
        CU(c, cudaGetLastError());
    }
    return HG_OK;
}

// the pixel kernel over the same descriptors, on the context stream
static int pw_fused_pixels(hg_ctx *c, const PwScratch &S, int nF, int max_ow, int max_oh, bool records_inline)
{
    const FusedFrame *frames = (const FusedFrame *)S.fframes->p;
    // rows per CTA = 16 * niter: long-lived CTAs amortise their start-up and keep the software pipeline full
    int niter = 16;
    while (niter > 1 && (long long)pwf_tiles_x(max_ow) * pwf_tiles_y(max_oh, niter) * nF < (long long)c->sm_count * 16) niter >>= 1;
    const int max_tiles = pwf_tiles_x(max_ow) * pwf_tiles_y(max_oh, niter);
    TRY(prof_begin(c));
    if (c->pwf_v1) pw_warp_fused_v1_kernel<<<dim3((unsigned)max_tiles, (unsigned)nF), PWF_THREADS, 0, c->stream>>>(frames, niter);
    else pw_warp_fused_kernel<<<dim3((unsigned)max_tiles, (unsigned)nF), PWF_THREADS, 0, c->stream>>>(frames, niter, records_inline ? 1 : 0);
    c->launches++;
    CU(c, cudaGetLastError());
    TRY(prof_end(c));
    return HG_OK;
}

// binning passes on `pre`, then the pixel kernel on the context stream; lane >= 0: the two are different streams, joined by
// the lane's events (ev_pre: bins ready; ev_pix: scratch free again)
static int pw_fused_launch(hg_ctx *c, const PwScratch &S, int lane, const float *dst_dev, int nF, int max_ow, int max_oh,
                           size_t max_bins, bool widths_aligned)
{
    const bool records_inline = widths_aligned && !c->pwf_v1 && c->pwf_records_inline && !pw_band_binning(c);
    cudaStream_t pre = lane >= 0 ? c->stream_pre : c->stream;
    TRY(pw_fused_prepass(c, S, pre, dst_dev, nF, max_ow, max_oh, max_bins, records_inline));
    if (lane >= 0) {
        CU(c, cudaEventRecord(c->ev_pre[lane], pre));
        CU(c, cudaStreamWaitEvent(c->stream, c->ev_pre[lane], 0));
    }
    return pw_fused_pixels(c, S, nF, max_ow, max_oh, records_inline);
}

// fused (map-free) inverse piecewise warp of nF frames in four launches; S.fstatus[f] != 0 afterwards means frame f could
// not be represented and must be redone with pw_inverse_general_frame.  lane < 0: everything on the context stream with
// scratch set 0; lane 0 / 1: binning on the high-priority stream with that lane's scratch (`lane_busy`: the lane's previous
// pixel kernel has to finish first)
static int pw_inverse_fused_chunk(hg_ctx *c, const float *dst_dev, const PwFrameHost *fr, int nF, int min_src_x,
                                  int min_src_y, int lane = -1, bool lane_busy = false)
{
    const size_t T = (size_t)c->n_tris;
    const PwScratch S = pw_scratch(c, lane < 0 ? 0 : lane);
    cudaStream_t pre = lane >= 0 ? c->stream_pre : c->stream;
    std::vector<FusedFrame> ff((size_t)nF);
    size_t total_bins = 0;
    for (int f = 0; f < nF; ++f) {
        const int bx = pwf_bins_x(fr[f].oW);
        total_bins += (size_t)bx * fr[f].oH;
    }
    int max_ow = 1, max_oh = 1;
    for (int f = 0; f < nF; ++f) {
        if (fr[f].oW > max_ow) max_ow = fr[f].oW;
        if (fr[f].oH > max_oh) max_oh = fr[f].oH;
    }
    TRY(pw_scratch_ensure(c, S, T, nF, total_bins));
    size_t bin0 = 0;
    for (int f = 0; f < nF; ++f) {
        FusedFrame &F = ff[(size_t)f];
        F.src = fr[f].src; F.out = fr[f].out;
        F.rec = (const TriRec *)S.rec->p + T * f;
        F.inv = (const float *)S.invd->p + 8 * T * f;
        F.yr = (const int2 *)S.yr->p + T * f;
        F.bin_cnt = (unsigned *)S.bin_cnt->p + 2 * bin0;
        F.bin_ent = (unsigned *)S.bin_ent->p + bin0 * PW_BIN_CAP;
        F.bin_run = (uint4 *)S.bin_run->p + 2 * bin0;
        F.status = (int *)S.fstatus->p + f;
        F.W = fr[f].W; F.H = fr[f].H;
        F.xOff = fr[f].xOff; F.yOff = fr[f].yOff; F.oW = fr[f].oW; F.oH = fr[f].oH;
        F.minSrcX = min_src_x; F.minSrcY = min_src_y;
        F.n_tris = c->n_tris;
        F.bins_x = pwf_bins_x(fr[f].oW);
        bin0 += (size_t)F.bins_x * fr[f].oH;
    }
    if (lane_busy) CU(c, cudaStreamWaitEvent(pre, c->ev_pix[lane], 0));
    CU(c, cudaMemcpyAsync(S.fframes->p, ff.data(), sizeof(FusedFrame) * (size_t)nF, cudaMemcpyHostToDevice, pre));
    if (!pw_band_binning(c)) CU(c, cudaMemsetAsync(S.bin_cnt->p, 0, sizeof(unsigned) * 2 * total_bins, pre));
    CU(c, cudaMemsetAsync(S.fstatus->p, 0, sizeof(int) * (size_t)nF, pre));
    size_t max_bins = 1;
    for (int f = 0; f < nF; ++f) {
        const size_t nb = (size_t)pwf_bins_x(fr[f].oW) * fr[f].oH;
        if (nb > max_bins) max_bins = nb;
    }
    bool aligned = true;
    for (int f = 0; f < nF; ++f) aligned = aligned && (fr[f].oW & 3) == 0;
    return pw_fused_launch(c, S, lane, dst_dev, nF, max_ow, max_oh, max_bins, aligned);
}

static bool pw_fused_possible(hg_ctx *c, const PwFrameHost &f)
{
    return !c->force_general && c->n_tris < PW_MAX_TRIS && f.oW <= 65536;
}

int hg_warp_piecewise_inverse(hg_ctx *c, const float *dst_pts, int x_off, int y_off, int o_w, int o_h,
                              int min_src_x, int min_src_y, uint8_t *out_host, void *out_dev)
{
    BIND(c);
    NEED(c, dst_pts, "dst_pts is NULL");
    if (!c->img) return fail(c, HG_ERR_STATE, "no image set (hg_image_set)");
    if (c->n_pts == 0) return fail(c, HG_ERR_STATE, "no mesh set (hg_piecewise_set_mesh)");
    TRY(check_window(c, x_off, y_off, o_w, o_h));
    if (min_src_x > (1 << 18) || min_src_x < -(1 << 18) || min_src_y > (1 << 18) || min_src_y < -(1 << 18))
        return fail(c, HG_ERR_UNSUPPORTED, "min_src (%d,%d) outside the supported range", min_src_x, min_src_y);
    TRY(check_points(c, dst_pts, c->n_pts, "dst_pts"));
    const size_t bytes = (size_t)o_w * o_h * 4;
    uint32_t *dst = nullptr;
    TRY(pick_out(c, out_dev, bytes, &dst));
    TRY(upload_dst_points(c, dst_pts, 1));
    // remember what the reference's shared map field would now hold (inverse map of these points), Q8
    c->last_inv_pts.assign(dst_pts, dst_pts + 2 * (size_t)c->n_pts);
    c->last_inv_mw = (double)o_w;
    c->last_inv_yoff = (double)y_off;
    c->last_inv_len = (long long)o_w * o_h;
    c->map32_current = false;
    PwFrameHost f{c->img, dst, c->W, c->H, x_off, y_off, o_w, o_h};
    bool general = !pw_fused_possible(c, f);
    if (!general) {
        TRY(pw_inverse_fused_chunk(c, (const float *)c->dst_pts.p, &f, 1, min_src_x, min_src_y));
        int *st = (int *)c->pinned;
        CU(c, cudaMemcpyAsync(st, c->fstatus.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        general = (*st != 0);
    }
    if (general) {
        TRY(pw_inverse_general_frame(c, (const float *)c->dst_pts.p, f, min_src_x, min_src_y));
        c->map32_current = true;
        c->n_general++;
    } else {
        c->n_fused++;
    }
    return finish_out(c, dst, bytes, out_host);
}

int hg_warp_piecewise_forward(hg_ctx *c, const float *dst_pts, int x_off, int y_off, int o_w, int o_h, int min_src_x,
                              int min_src_y, int max_src_x, int max_src_y, int use_inverse_map, uint8_t *out_host,
                              void *out_dev)
{
    BIND(c);
    NEED(c, dst_pts, "dst_pts is NULL");
    if (!c->img) return fail(c, HG_ERR_STATE, "no image set (hg_image_set)");
    if (c->n_pts == 0) return fail(c, HG_ERR_STATE, "no mesh set (hg_piecewise_set_mesh)");
    TRY(check_window(c, x_off, y_off, o_w, o_h));
    const long long dom_w = (long long)max_src_x - min_src_x, dom_h = (long long)max_src_y - min_src_y;
    if (min_src_x > (1 << 18) || min_src_x < -(1 << 18) || min_src_y > (1 << 18) || min_src_y < -(1 << 18) ||
        dom_w > 65536 || dom_h > 65536 || dom_w * dom_h >= (1LL << 31))
        return fail(c, HG_ERR_UNSUPPORTED, "source-point bounding box outside the supported range");
    TRY(check_points(c, dst_pts, c->n_pts, "dst_pts"));
    const size_t bytes = (size_t)o_w * o_h * 4;
    uint32_t *dst = nullptr;
    TRY(pick_out(c, out_dev, bytes, &dst));
    TRY(upload_dst_points(c, dst_pts, 1));
    // forward matrices from (src, dst); edge equations of the SOURCE triangles for the forward map (H.js:817-832)
    TRY(launch_setup(c, (const float *)c->dst_pts.p, (const float *)c->src_pts.p, nullptr, nullptr));
    if (!use_inverse_map) {
        const long long len = dom_w > 0 && dom_h > 0 ? dom_w * dom_h : 0;
        TRY(launch_fill(c, (double)dom_w, (double)min_src_y, len));
        c->map32_current = false;
    } else if (!c->map32_current) {
        // the map the reference would still hold (inverse map of the last inverse warp) was never materialised by
        // the fused path: rebuild it now from the remembered points (H.js:957 aliasing, Q8)
        if (c->last_inv_len < 0) return fail(c, HG_ERR_STATE, "use_inverse_map without a previous inverse warp");
        TRY(upload_dst_points(c, c->last_inv_pts.data(), 1));
        TRY(launch_setup(c, nullptr, (const float *)c->dst_pts.p, nullptr, nullptr));
        TRY(launch_fill(c, c->last_inv_mw, c->last_inv_yoff, c->last_inv_len));
        c->map32_current = true;
        TRY(upload_dst_points(c, dst_pts, 1));
        TRY(launch_setup(c, (const float *)c->dst_pts.p, (const float *)c->src_pts.p, nullptr, nullptr));
    }
    FwdArgs a{};
    a.src = c->img;
    a.map32 = (const int *)c->map32.p;
    a.map_len = c->map_len;
    a.rec = (const TriRec *)c->rec.p;
    a.n_tris = c->n_tris;
    a.W = c->W; a.H = c->H;
    a.xOff = x_off; a.yOff = y_off; a.oW = o_w; a.oH = o_h;
    a.minX = min_src_x; a.minY = min_src_y;
    a.domW = dom_w > 0 ? (int)dom_w : 0;
    a.domH = dom_h > 0 ? (int)dom_h : 0;
    return run_forward(c, a, true, dst, bytes, out_host);
}

int hg_warp_piecewise_inverse_batch(hg_ctx *c, const float *dst_pts, const hg_frame *frames, int n_frames,
                                    int min_src_x, int min_src_y)
{
    BIND(c);
    NEED(c, dst_pts && frames, "NULL argument");
    NEED(c, n_frames >= 1, "n_frames must be >= 1");
    if (c->n_pts == 0) return fail(c, HG_ERR_STATE, "no mesh set (hg_piecewise_set_mesh)");
    if (min_src_x > (1 << 18) || min_src_x < -(1 << 18) || min_src_y > (1 << 18) || min_src_y < -(1 << 18))
        return fail(c, HG_ERR_UNSUPPORTED, "min_src (%d,%d) outside the supported range", min_src_x, min_src_y);
    std::vector<PwFrameHost> fr((size_t)n_frames);
    for (int f = 0; f < n_frames; ++f) {
        const hg_frame &h = frames[f];
        TRY(check_window(c, h.x_off, h.y_off, h.o_w, h.o_h));
        NEED(c, h.out_dev && ((uintptr_t)h.out_dev & 15) == 0, "frame out_dev must be a 16-byte aligned device pointer");
        PwFrameHost &g = fr[(size_t)f];
        if (h.src_dev) {
            TRY(check_image_dims(c, h.src_w, h.src_h));
            g.src = (const uint32_t *)h.src_dev; g.W = h.src_w; g.H = h.src_h;
        } else {
            if (!c->img) return fail(c, HG_ERR_STATE, "no image set (hg_image_set)");
            g.src = c->img; g.W = c->W; g.H = c->H;
        }
        g.out = (uint32_t *)h.out_dev;
        g.xOff = h.x_off; g.yOff = h.y_off; g.oW = h.o_w; g.oH = h.o_h;
    }
    const size_t pts_per_frame = 2 * (size_t)c->n_pts;
    TRY(check_point_floats(c, dst_pts, pts_per_frame * (size_t)n_frames, "dst_pts"));
    TRY(ensure(c, c->dst_pts, sizeof(float) * pts_per_frame * (size_t)n_frames));
    TRY(ensure_pinned(c, sizeof(int) * (size_t)n_frames));
    c->map32_current = false;
    c->last_inv_len = -1;
    // chunk so that the per-frame scratch (triangle records, inverse matrices, bins) stays below ~1.5 GB
    size_t per_frame = (sizeof(TriRec) + 48) * (size_t)c->n_tris;
    for (int f = 0; f < n_frames; ++f) {
        const size_t b = (size_t)pwf_bins_x(fr[(size_t)f].oW) * fr[(size_t)f].oH * (4 + 4 * PW_BIN_CAP + 32);
        if (b + (sizeof(TriRec) + 48) * (size_t)c->n_tris > per_frame) per_frame = b + (sizeof(TriRec) + 48) * (size_t)c->n_tris;
    }
    int chunk = (int)((1500ull << 20) / (per_frame ? per_frame : 1));
    if (chunk < 1) chunk = 1;
    // ... and at most 128 frames (64 with two lanes), so that a larger batch runs as a pipeline: the host stages the destiny
    // points and frame descriptors of chunk k+1 while chunk k computes (the status flags come back through pinned memory, so
    // nothing below blocks before the one synchronisation at the end).  Measured on 4K frames of the 162-triangle mesh, whole
    // step against the HBM bound: chunks of 16 / 32 / 64 frames 0.408 / 0.431 / 0.441 (64 frames per call, earlier kernels);
    // 64 / 128 frames 0.504 / 0.518 (128 frames per call) — launches this size are worth more than the overlap of a 2 MB upload.
    const bool fused = pw_fused_possible(c, fr[0]);
    // One lane or two: pw_two_lanes().  HG_PW_LANES=0 / 1 forces either.
    const bool two = pw_two_lanes(c);
    const int want = c->pw_chunk > 0 ? c->pw_chunk : (two ? 64 : 128);
    if (chunk > want) chunk = want;
    const bool lanes = fused && two && n_frames > chunk;
    int *status = (int *)c->pin_big;
    if (lanes) {
        // both lanes sized for the largest chunk before anything is in flight
        size_t most_bins = 0;
        for (int f0 = 0; f0 < n_frames; f0 += chunk) {
            size_t b = 0;
            for (int f = f0; f < n_frames && f < f0 + chunk; ++f) b += (size_t)pwf_bins_x(fr[(size_t)f].oW) * fr[(size_t)f].oH;
            if (b > most_bins) most_bins = b;
        }
        for (int l = 0; l < 2; ++l) TRY(pw_scratch_ensure(c, pw_scratch(c, l), (size_t)c->n_tris, chunk, most_bins));
        TRY(pw_lanes_begin(c));
    }
    int k = 0;
    for (int f0 = 0; f0 < n_frames; f0 += chunk, ++k) {
        const int nf = n_frames - f0 < chunk ? n_frames - f0 : chunk;
        float *dd = (float *)c->dst_pts.p + pts_per_frame * (size_t)f0;
        // the points go up in front of the lane's wait for its scratch: the host never blocks behind a pixel kernel
        CU(c, cudaMemcpyAsync(dd, dst_pts + pts_per_frame * (size_t)f0, sizeof(float) * pts_per_frame * (size_t)nf,
                              cudaMemcpyHostToDevice, lanes ? c->stream_pre : c->stream));
        if (fused) {
            TRY(pw_inverse_fused_chunk(c, dd, fr.data() + f0, nf, min_src_x, min_src_y, lanes ? (k & 1) : -1, k >= 2));
            // the scratch is reused two chunks on: collect this chunk's status first (stream-ordered copy)
            CU(c, cudaMemcpyAsync(status + f0, pw_scratch(c, lanes ? (k & 1) : 0).fstatus->p, sizeof(int) * (size_t)nf,
                                  cudaMemcpyDeviceToHost, c->stream));
            if (lanes) CU(c, cudaEventRecord(c->ev_pix[k & 1], c->stream));
        } else {
            for (int f = 0; f < nf; ++f) status[f0 + f] = 1;
        }
    }
    CU(c, cudaStreamSynchronize(c->stream));
    for (int f = 0; f < n_frames; ++f) {
        if (status[(size_t)f]) {  // not representable by the bins: the general, map-based path (exact for everything)
            TRY(pw_inverse_general_frame(c, (const float *)c->dst_pts.p + pts_per_frame * (size_t)f, fr[(size_t)f],
                                         min_src_x, min_src_y));
            c->n_general++;
        } else {
            c->n_fused++;
        }
    }
    return HG_OK;
}

int hg_debug_piecewise_stats(hg_ctx *c, uint64_t *frames_fused, uint64_t *frames_general)
{
    if (!c || !frames_fused || !frames_general) return HG_ERR_INVALID;
    *frames_fused = c->n_fused;
    *frames_general = c->n_general;
    return HG_OK;
}

int hg_debug_piecewise_binning(hg_ctx *c, int mode)
{
    if (!c || mode < 0 || mode > 2) return HG_ERR_INVALID;
    c->pw_binning = mode;
    return HG_OK;
}

int hg_debug_force_general(hg_ctx *c, int on)
{
    if (!c) return HG_ERR_INVALID;
    c->force_general = on != 0;
    return HG_OK;
}

/* ------------------------------------------------------------------ pipelined host-to-host stream */
struct hg_pipe_slot {
    void *d_src = nullptr, *d_out = nullptr;
    char *d_small = nullptr;  // 256 B device scratch: [0,64) dst pts, [64,128) src pts, [128,192) matrix
    char *h_small = nullptr;  // 128 B pinned staging for the points
    cudaEvent_t in_done = nullptr, k_done = nullptr, out_done = nullptr;
    bool busy = false;
    CUtensorMap tm[GEO_NBOX];  // over d_src
    bool tm_ok = false;
};

struct hg_pipe {
    hg_ctx *c = nullptr;
    bool piecewise = false;
    int kind = 0, W = 0, H = 0, max_ow = 0, max_oh = 0, depth = 0;
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    std::vector<hg_pipe_slot> slots;
    uint64_t next = 0;
};

int hg_pipe_create(hg_ctx *c, int kind, int src_w, int src_h, int max_out_w, int max_out_h, int depth, hg_pipe **out)
{
    BIND(c);
    NEED(c, out, "out is NULL");
    *out = nullptr;
    NEED(c, kind == HG_AFFINE || kind == HG_PROJECTIVE, "kind must be HG_AFFINE or HG_PROJECTIVE");
    NEED(c, depth >= 1 && depth <= 16, "depth must be in [1,16]");
    TRY(check_image_dims(c, src_w, src_h));
    TRY(check_window(c, 0, 0, max_out_w, max_out_h));
    hg_pipe *p = new hg_pipe();
    p->c = c;
    p->kind = kind;
    p->W = src_w;
    p->H = src_h;
    p->max_ow = max_out_w;
    p->max_oh = max_out_h;
    p->depth = depth;
    p->slots.resize((size_t)depth);
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess && r != cudaSuccess) e = r; return r == cudaSuccess; };
    ok(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
    ok(cudaStreamCreateWithFlags(&p->s_k, cudaStreamNonBlocking));
    ok(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
    for (auto &sl : p->slots) {
        ok(cudaMalloc(&sl.d_src, (size_t)src_w * src_h * 4));
        ok(cudaMalloc(&sl.d_out, (size_t)max_out_w * max_out_h * 4));
        ok(cudaMalloc((void **)&sl.d_small, 256));
        ok(cudaHostAlloc((void **)&sl.h_small, 128, cudaHostAllocDefault));
        ok(cudaEventCreateWithFlags(&sl.in_done, cudaEventDisableTiming));
        ok(cudaEventCreateWithFlags(&sl.k_done, cudaEventDisableTiming));
        ok(cudaEventCreateWithFlags(&sl.out_done, cudaEventDisableTiming));
        sl.tm_ok = sl.d_src && encode_tmaps(c, sl.d_src, src_w, src_h, sl.tm);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        hg_pipe_destroy(p);
        return fail(c, e == cudaErrorMemoryAllocation ? HG_ERR_NOMEM : HG_ERR_CUDA, "hg_pipe_create: %s", cudaGetErrorString(e));
    }
    *out = p;
    return HG_OK;
}

int hg_pipe_destroy(hg_pipe *p)
{
    if (!p) return HG_ERR_INVALID;
    cudaSetDevice(p->c->device);
    if (p->s_in) cudaStreamSynchronize(p->s_in);
    if (p->s_k) cudaStreamSynchronize(p->s_k);
    if (p->s_out) cudaStreamSynchronize(p->s_out);
    for (auto &sl : p->slots) {
        if (sl.d_src) cudaFree(sl.d_src);
        if (sl.d_out) cudaFree(sl.d_out);
        if (sl.d_small) cudaFree(sl.d_small);
        if (sl.h_small) cudaFreeHost(sl.h_small);
        if (sl.in_done) cudaEventDestroy(sl.in_done);
        if (sl.k_done) cudaEventDestroy(sl.k_done);
        if (sl.out_done) cudaEventDestroy(sl.out_done);
    }
    if (p->s_in) cudaStreamDestroy(p->s_in);
    if (p->s_k) cudaStreamDestroy(p->s_k);
    if (p->s_out) cudaStreamDestroy(p->s_out);
    delete p;
    return HG_OK;
}

int hg_pipe_submit(hg_pipe *p, const uint8_t *rgba_host, const double *dst_pts, const double *src_pts, int x_off,
                   int y_off, int o_w, int o_h, uint8_t *out_host, uint64_t *ticket)
{
    if (!p) return HG_ERR_INVALID;
    hg_ctx *c = p->c;
    BIND(c);
    NEED(c, rgba_host && dst_pts && src_pts && out_host, "NULL argument");
    NEED(c, !p->piecewise, "a piecewise pipe takes hg_pipe_submit_piecewise");
    TRY(check_window(c, x_off, y_off, o_w, o_h));
    if (o_w > p->max_ow || o_h > p->max_oh || (long long)o_w * o_h > (long long)p->max_ow * p->max_oh)
        return fail(c, HG_ERR_INVALID, "output %dx%d exceeds the pipe's maximum %dx%d", o_w, o_h, p->max_ow, p->max_oh);
    hg_pipe_slot &sl = p->slots[(size_t)(p->next % (uint64_t)p->depth)];
    if (sl.busy) CU(c, cudaEventSynchronize(sl.out_done));  // the slot's previous frame has left the device
    const size_t pb = p->kind == HG_AFFINE ? 48 : 64;
    SolvePoints sp{};
    memcpy(sp.src, dst_pts, pb);  // inverse matrix = calculateTransformMatrix(kind, dst, src) (H.js:994)
    memcpy(sp.dst, src_pts, pb);
    // copy-in stream: the image
    CU(c, cudaMemcpyAsync(sl.d_src, rgba_host, (size_t)p->W * p->H * 4, cudaMemcpyHostToDevice, p->s_in));
    CU(c, cudaEventRecord(sl.in_done, p->s_in));
    // compute stream: solve (points travel as kernel parameters) + pixel loop
    CU(c, cudaStreamWaitEvent(p->s_k, sl.in_done, 0));
    solve_points_kernel<<<1, 32, 0, p->s_k>>>(sp, p->kind == HG_AFFINE ? 0 : 1, (float *)(sl.d_small + 128),
                                              (double *)(sl.d_small + 128));
    GeoParams P{};
    P.one.src = (const uint32_t *)sl.d_src;
    P.one.out = (uint32_t *)sl.d_out;
    P.one.W = p->W;
    P.one.H = p->H;
    P.one.xOff = x_off;
    P.one.yOff = y_off;
    P.one.oW = o_w;
    P.one.oH = o_h;
    P.many = nullptr;
    P.mats_dev = sl.d_small + 128;
    P.has_tm = sl.tm_ok ? 1 : 0;
    if (sl.tm_ok) memcpy(P.tm_val, sl.tm, sizeof sl.tm);
    c->launches++;
    CU(c, cudaGetLastError());
    const int sampling = c->sampling;
    c->sampling = HG_NEAREST;  // the pipe is the reference's nearest-neighbour path
    const int lr = launch_geo(c, p->kind, P, o_w, o_h, 1, sl.tm_ok, p->s_k, points_map_is_rotated(dst_pts, src_pts));
    c->sampling = sampling;
    TRY(lr);
    CU(c, cudaEventRecord(sl.k_done, p->s_k));
    // copy-out stream
    CU(c, cudaStreamWaitEvent(p->s_out, sl.k_done, 0));
    CU(c, cudaMemcpyAsync(out_host, sl.d_out, (size_t)o_w * o_h * 4, cudaMemcpyDeviceToHost, p->s_out));
    CU(c, cudaEventRecord(sl.out_done, p->s_out));
    sl.busy = true;
    if (ticket) *ticket = p->next;
    p->next++;
    return HG_OK;
}

int hg_pipe_create_piecewise(hg_ctx *c, int src_w, int src_h, int max_out_w, int max_out_h, int depth, hg_pipe **out)
{
    const int r = hg_pipe_create(c, HG_AFFINE, src_w, src_h, max_out_w, max_out_h, depth, out);
    if (r == HG_OK) (*out)->piecewise = true;
    return r;
}

namespace {
// the fused / general piecewise helpers enqueue on c->stream: run them on another stream for the length of a scope
struct StreamSwap {
    hg_ctx *c;
    cudaStream_t saved;
    StreamSwap(hg_ctx *ctx, cudaStream_t s) : c(ctx), saved(ctx->stream) { c->stream = s; }
    ~StreamSwap() { c->stream = saved; }
};
}  // namespace

int hg_pipe_submit_piecewise(hg_pipe *p, const uint8_t *rgba_host, const float *dst_pts, int min_src_x, int min_src_y,
                             uint8_t *out_host, int32_t window_out[4], uint64_t *ticket)
{
    if (!p) return HG_ERR_INVALID;
    hg_ctx *c = p->c;
    BIND(c);
    NEED(c, p->piecewise, "not a piecewise pipe (hg_pipe_create_piecewise)");
    NEED(c, dst_pts && out_host && window_out, "NULL argument");
    if (c->n_pts == 0) return fail(c, HG_ERR_STATE, "no mesh set (hg_piecewise_set_mesh)");
    if (min_src_x > (1 << 18) || min_src_x < -(1 << 18) || min_src_y > (1 << 18) || min_src_y < -(1 << 18))
        return fail(c, HG_ERR_UNSUPPORTED, "min_src (%d,%d) outside the supported range", min_src_x, min_src_y);
    TRY(check_points(c, dst_pts, c->n_pts, "dst_pts"));
    // output window on the host, the arithmetic of H.js:706-710 + minmaxXYofArray H.js:1558 (floats compare exactly;
    // Math.round of a float32 value is exact in double): the copy-out below needs its size before the frame runs
    float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
    for (int i = 0; i < c->n_pts; ++i) {
        const float x = dst_pts[2 * i], y = dst_pts[2 * i + 1];
        if (x > mxx) mxx = x;
        if (x < mnx) mnx = x;
        if (y > mxy) mxy = y;
        if (y < mny) mny = y;
    }
    auto jsr = [](double v) { const double fl = std::floor(v); return (v - fl) >= 0.5 ? fl + 1.0 : fl; };
    const double x0 = jsr((double)mnx), y0 = jsr((double)mny), w = jsr((double)mxx) - x0, h = jsr((double)mxy) - y0;
    if (!(w >= 1.0 && h >= 1.0 && w <= (double)p->max_ow && h <= (double)p->max_oh && std::fabs(x0) <= 262144.0 &&
          std::fabs(y0) <= 262144.0))
        return fail(c, HG_ERR_UNSUPPORTED, "output window %gx%g at (%g,%g) is empty or exceeds the pipe's maximum %dx%d", w, h, x0,
                    y0, p->max_ow, p->max_oh);
    const int xo = (int)x0, yo = (int)y0, oW = (int)w, oH = (int)h;
    if (!rgba_host && !c->img) return fail(c, HG_ERR_STATE, "no image set (hg_image_set) and no frame image given");
    hg_pipe_slot &sl = p->slots[(size_t)(p->next % (uint64_t)p->depth)];
    if (sl.busy) CU(c, cudaEventSynchronize(sl.out_done));  // the slot's previous frame has left the device
    const uint32_t *src = c->img;
    int W = c->W, H = c->H;
    if (rgba_host) {
        CU(c, cudaMemcpyAsync(sl.d_src, rgba_host, (size_t)p->W * p->H * 4, cudaMemcpyHostToDevice, p->s_in));
        CU(c, cudaEventRecord(sl.in_done, p->s_in));
        CU(c, cudaStreamWaitEvent(p->s_k, sl.in_done, 0));
        src = (const uint32_t *)sl.d_src;
        W = p->W;
        H = p->H;
    }
    {
        StreamSwap swap(c, p->s_k);  // the context's piecewise scratch is used in s_k order by every frame of the pipe
        TRY(upload_dst_points(c, dst_pts, 1));
        c->map32_current = false;
        c->last_inv_len = -1;
        PwFrameHost f{src, (uint32_t *)sl.d_out, W, H, xo, yo, oW, oH};
        bool general = !pw_fused_possible(c, f);
        if (!general) {
            TRY(pw_inverse_fused_chunk(c, (const float *)c->dst_pts.p, &f, 1, min_src_x, min_src_y));
            int *st = (int *)c->pinned;
            CU(c, cudaMemcpyAsync(st, c->fstatus.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            CU(c, cudaStreamSynchronize(c->stream));  // compute stream only: the previous frame's copy-out keeps running
            general = (*st != 0);
        }
        if (general) {
            TRY(pw_inverse_general_frame(c, (const float *)c->dst_pts.p, f, min_src_x, min_src_y));
            c->n_general++;
        } else {
            c->n_fused++;
        }
    }
    CU(c, cudaEventRecord(sl.k_done, p->s_k));
    CU(c, cudaStreamWaitEvent(p->s_out, sl.k_done, 0));
    CU(c, cudaMemcpyAsync(out_host, sl.d_out, (size_t)oW * oH * 4, cudaMemcpyDeviceToHost, p->s_out));
    CU(c, cudaEventRecord(sl.out_done, p->s_out));
    sl.busy = true;
    window_out[0] = xo; window_out[1] = yo; window_out[2] = oW; window_out[3] = oH;
    if (ticket) *ticket = p->next;
    p->next++;
    return HG_OK;
}

int hg_pipe_wait(hg_pipe *p, uint64_t ticket)
{
    if (!p) return HG_ERR_INVALID;
    hg_ctx *c = p->c;
    BIND(c);
    NEED(c, ticket < p->next, "unknown ticket");
    if (ticket + (uint64_t)p->depth < p->next) return HG_OK;  // its slot was already recycled, i.e. completed
    hg_pipe_slot &sl = p->slots[(size_t)(ticket % (uint64_t)p->depth)];
    if (sl.busy) CU(c, cudaEventSynchronize(sl.out_done));
    return HG_OK;
}

int hg_pipe_flush(hg_pipe *p)
{
    if (!p) return HG_ERR_INVALID;
    hg_ctx *c = p->c;
    BIND(c);
    CU(c, cudaStreamSynchronize(p->s_out));
    CU(c, cudaStreamSynchronize(p->s_k));
    CU(c, cudaStreamSynchronize(p->s_in));
    for (auto &sl : p->slots) sl.busy = false;
    return HG_OK;
}

/* ------------------------------------------------------------------ batches: per-frame solve, forward loops */
int hg_warp_inverse_points_batch(hg_ctx *c, int kind, const double *dst_pts, const double *src_pts, const hg_frame *frames,
                                 int n_frames)
{
    BIND(c);
    NEED(c, dst_pts && src_pts && frames, "NULL argument");
    NEED(c, kind == HG_AFFINE || kind == HG_PROJECTIVE, "kind must be HG_AFFINE or HG_PROJECTIVE");
    NEED(c, n_frames >= 1 && n_frames <= 32768, "n_frames must be in [1, 32768]");
    std::vector<GeoFrame> gf((size_t)n_frames);
    int max_ow = 1, max_oh = 1;
    for (int f = 0; f < n_frames; ++f) {
        const hg_frame &h = frames[f];
        TRY(check_window(c, h.x_off, h.y_off, h.o_w, h.o_h));
        NEED(c, h.out_dev && ((uintptr_t)h.out_dev & 15) == 0, "frame out_dev must be a 16-byte aligned device pointer");
        GeoFrame &g = gf[(size_t)f];
        if (h.src_dev) {
            TRY(check_image_dims(c, h.src_w, h.src_h));
            g.src = (const uint32_t *)h.src_dev; g.W = h.src_w; g.H = h.src_h;
        } else {
            if (!c->img) return fail(c, HG_ERR_STATE, "no image set (hg_image_set)");
            g.src = c->img; g.W = c->W; g.H = c->H;
        }
        g.tm = nullptr;  // direct-gather kernel (the staged kernel is an opt-in experiment of the matrix entry points)
        g.out = (uint32_t *)h.out_dev;
        g.xOff = h.x_off; g.yOff = h.y_off; g.oW = h.o_w; g.oH = h.o_h;
        if (h.o_w > max_ow) max_ow = h.o_w;
        if (h.o_h > max_oh) max_oh = h.o_h;
    }
    const size_t pb = kind == HG_AFFINE ? 48 : 64, mstride = kind == HG_AFFINE ? 24 : 64;
    TRY(ensure(c, c->frames, sizeof(GeoFrame) * (size_t)n_frames));
    TRY(ensure(c, c->mats, mstride * (size_t)n_frames));
    TRY(ensure(c, c->pts_batch, 2 * pb * (size_t)n_frames));
    char *pd = (char *)c->pts_batch.p;
    CU(c, cudaMemcpyAsync(c->frames.p, gf.data(), sizeof(GeoFrame) * (size_t)n_frames, cudaMemcpyHostToDevice, c->stream));
    // inverse matrix of frame f = calculateTransformMatrix(kind, dstPoints_f, srcPoints_f)  (H.js:994)
    CU(c, cudaMemcpyAsync(pd, dst_pts, pb * (size_t)n_frames, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(pd + pb * (size_t)n_frames, src_pts, pb * (size_t)n_frames, cudaMemcpyHostToDevice, c->stream));
    SolveArgs a{};
    a.src = (const double *)pd;
    a.dst = (const double *)(pd + pb * (size_t)n_frames);
    a.out_f = (float *)c->mats.p;
    a.out_d = (double *)c->mats.p;
    a.n = n_frames;
    a.op = kind == HG_AFFINE ? 0 : 1;
    TRY(launch_solve(c, a));
    GeoParams P{};
    P.many = (const GeoFrame *)c->frames.p;
    P.mats_dev = c->mats.p;
    return launch_geo(c, kind, P, max_ow, max_oh, n_frames, false, c->stream, points_map_is_rotated(dst_pts, src_pts));
}

static int fill_fwd_frame(hg_ctx *c, const hg_frame &h, FwdArgs &a)
{
    TRY(check_window(c, h.x_off, h.y_off, h.o_w, h.o_h));
    NEED(c, h.out_dev && ((uintptr_t)h.out_dev & 15) == 0, "frame out_dev must be a 16-byte aligned device pointer");
    if (h.src_dev) {
        TRY(check_image_dims(c, h.src_w, h.src_h));
        a.src = (const uint32_t *)h.src_dev; a.W = h.src_w; a.H = h.src_h;
    } else {
        if (!c->img) return fail(c, HG_ERR_STATE, "no image set (hg_image_set)");
        a.src = c->img; a.W = c->W; a.H = c->H;
    }
    a.out = (uint32_t *)h.out_dev;
    a.xOff = h.x_off; a.yOff = h.y_off; a.oW = h.o_w; a.oH = h.o_h;
    return HG_OK;
}

int hg_warp_forward_batch(hg_ctx *c, int kind, const void *fwd_matrices, const hg_frame *frames, int n_frames)
{
    BIND(c);
    NEED(c, fwd_matrices && frames, "NULL argument");
    NEED(c, kind == HG_AFFINE || kind == HG_PROJECTIVE, "kind must be HG_AFFINE or HG_PROJECTIVE");
    NEED(c, n_frames >= 1, "n_frames must be >= 1");
    std::vector<FwdArgs> fa((size_t)n_frames);
    for (int f = 0; f < n_frames; ++f) {
        FwdArgs &a = fa[(size_t)f];
        a = FwdArgs{};
        TRY(fill_fwd_frame(c, frames[f], a));
        a.kind = kind;
        if (kind == HG_AFFINE)
            for (int k = 0; k < 6; ++k) a.mat[k] = (double)((const float *)fwd_matrices)[6 * (size_t)f + k];
        else
            for (int k = 0; k < 8; ++k) a.mat[k] = ((const double *)fwd_matrices)[8 * (size_t)f + k];
        a.minX = 0; a.minY = 0; a.domW = a.W; a.domH = a.H;  // for (y < H) for (x < W), H.js:919-920
    }
    return run_forward_batch(c, fa, false);
}

int hg_warp_piecewise_forward_batch(hg_ctx *c, const float *dst_pts, const hg_frame *frames, int n_frames, int min_src_x,
                                    int min_src_y, int max_src_x, int max_src_y)
{
    BIND(c);
    NEED(c, dst_pts && frames, "NULL argument");
    NEED(c, n_frames >= 1, "n_frames must be >= 1");
    if (c->n_pts == 0) return fail(c, HG_ERR_STATE, "no mesh set (hg_piecewise_set_mesh)");
    const long long dom_w = (long long)max_src_x - min_src_x, dom_h = (long long)max_src_y - min_src_y;
    if (min_src_x > (1 << 18) || min_src_x < -(1 << 18) || min_src_y > (1 << 18) || min_src_y < -(1 << 18) ||
        dom_w > 65536 || dom_h > 65536 || dom_w * dom_h >= (1LL << 31))
        return fail(c, HG_ERR_UNSUPPORTED, "source-point bounding box outside the supported range");
    const size_t T = (size_t)c->n_tris, pts_per_frame = 2 * (size_t)c->n_pts;
    TRY(check_point_floats(c, dst_pts, pts_per_frame * (size_t)n_frames, "dst_pts"));
    TRY(upload_dst_points(c, dst_pts, (size_t)n_frames));
    c->map32_current = false;
    c->last_inv_len = -1;
    // chunks: the per-frame triangle records (forward matrices) stay below ~1 GB
    int chunk = (int)((1000ull << 20) / (sizeof(TriRec) * (T ? T : 1)));
    if (chunk < 1) chunk = 1;
    if (chunk > 4096) chunk = 4096;
    const long long len = dom_w > 0 && dom_h > 0 ? dom_w * dom_h : 0;
    for (int f0 = 0; f0 < n_frames; f0 += chunk) {
        const int nf = n_frames - f0 < chunk ? n_frames - f0 : chunk;
        TRY(ensure(c, c->rec, sizeof(TriRec) * (T ? T : 1) * (size_t)nf));
        if (T > 0) {
            // forward matrices from (src, dst_f); edge equations of the SOURCE triangles for the forward map (H.js:817-832)
            PwSetupArgs s{};
            s.src_pts = (const float *)c->src_pts.p;
            s.dst_pts = (const float *)c->dst_pts.p + pts_per_frame * (size_t)f0;
            s.map_pts = (const float *)c->src_pts.p;
            s.tris = (const uint32_t *)c->tris.p;
            s.rec = (TriRec *)c->rec.p;
            s.n_tris = c->n_tris;
            s.dst_stride = pts_per_frame;
            s.rec_stride = T;
            pw_setup_kernel<<<dim3((unsigned)((T + 127) / 128), (unsigned)nf), 128, 0, c->stream>>>(s);
            c->launches++;
            CU(c, cudaGetLastError());
        }
        // the forward map depends on the source points only: built once per call from the first chunk's records (the
        // reference builds it once per source-point set, H.js:759)
        if (f0 == 0) TRY(launch_fill(c, (double)dom_w, (double)min_src_y, len));
        std::vector<FwdArgs> fa((size_t)nf);
        for (int f = 0; f < nf; ++f) {
            FwdArgs &a = fa[(size_t)f];
            a = FwdArgs{};
            TRY(fill_fwd_frame(c, frames[f0 + f], a));
            a.map32 = (const int *)c->map32.p;
            a.map_len = c->map_len;
            a.rec = (const TriRec *)c->rec.p + T * (size_t)f;
            a.n_tris = c->n_tris;
            a.minX = min_src_x; a.minY = min_src_y;
            a.domW = dom_w > 0 ? (int)dom_w : 0;
            a.domH = dom_h > 0 ? (int)dom_h : 0;
        }
        TRY(run_forward_batch(c, fa, true));
    }
    return HG_OK;
}

/* ------------------------------------------------------------------ checksums */
int hg_checksum_frames(hg_ctx *c, const hg_frame *frames, int n_frames, uint64_t *out_host)
{
    BIND(c);
    NEED(c, frames && out_host, "NULL argument");
    NEED(c, n_frames >= 1 && n_frames <= 65535, "n_frames must be in [1, 65535]");
    std::vector<ChecksumFrame> cf((size_t)n_frames);
    long long max_n = 1;
    for (int f = 0; f < n_frames; ++f) {
        NEED(c, frames[f].out_dev && frames[f].o_w >= 0 && frames[f].o_h >= 0, "bad frame");
        cf[(size_t)f].px = (const uint32_t *)frames[f].out_dev;
        cf[(size_t)f].n = (long long)frames[f].o_w * frames[f].o_h;
        if (cf[(size_t)f].n > max_n) max_n = cf[(size_t)f].n;
    }
    TRY(ensure(c, c->cs_frames, sizeof(ChecksumFrame) * (size_t)n_frames));
    TRY(ensure(c, c->cs_out, sizeof(uint64_t) * (size_t)n_frames));
    TRY(ensure_pinned(c, sizeof(uint64_t) * (size_t)n_frames));
    CU(c, cudaMemcpyAsync(c->cs_frames.p, cf.data(), sizeof(ChecksumFrame) * (size_t)n_frames, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemsetAsync(c->cs_out.p, 0, sizeof(uint64_t) * (size_t)n_frames, c->stream));
    long long blocks = (max_n + 256 * 8 - 1) / (256 * 8);
    const long long cap = (long long)c->sm_count * 8 / (n_frames < 8 ? 1 : 4);
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    checksum_frames_kernel<<<dim3((unsigned)blocks, (unsigned)n_frames), 256, 0, c->stream>>>((const ChecksumFrame *)c->cs_frames.p,
                                                                                            (unsigned long long *)c->cs_out.p);
    c->launches++;
    CU(c, cudaGetLastError());
    CU(c, cudaMemcpyAsync(c->pin_big, c->cs_out.p, sizeof(uint64_t) * (size_t)n_frames, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    memcpy(out_host, c->pin_big, sizeof(uint64_t) * (size_t)n_frames);
    return HG_OK;
}

/* ------------------------------------------------------------------ streamed piecewise frames */
size_t hg_stream_slot_bytes(int max_out_w, int max_out_h)
{
    if (max_out_w < 1 || max_out_h < 1) return 0;
    return (((size_t)max_out_w * (size_t)max_out_h * 4) + 255) & ~(size_t)255;
}

int hg_warp_piecewise_stream(hg_ctx *c, const float *dst_pts, int n_frames, int64_t first_frame, int min_src_x, int min_src_y,
                             const void *src_ring_dev, int n_src, int src_w, int src_h, void *out_ring_dev, int n_slots,
                             int max_out_w, int max_out_h, hg_stream_info *info_out)
{
    BIND(c);
    NEED(c, dst_pts && out_ring_dev, "NULL argument");
    NEED(c, n_frames >= 1 && first_frame >= 0, "n_frames must be >= 1 and first_frame >= 0");
    NEED(c, n_slots >= 1, "n_slots must be >= 1");
    NEED(c, ((uintptr_t)out_ring_dev & 255) == 0, "out_ring_dev must be 256-byte aligned");
    if (c->n_pts == 0) return fail(c, HG_ERR_STATE, "no mesh set (hg_piecewise_set_mesh)");
    if (min_src_x > (1 << 18) || min_src_x < -(1 << 18) || min_src_y > (1 << 18) || min_src_y < -(1 << 18))
        return fail(c, HG_ERR_UNSUPPORTED, "min_src (%d,%d) outside the supported range", min_src_x, min_src_y);
    TRY(check_window(c, 0, 0, max_out_w, max_out_h));
    const uint32_t *src = nullptr;
    int W = 0, H = 0;
    if (src_ring_dev) {
        NEED(c, n_src >= 1 && ((uintptr_t)src_ring_dev & 3) == 0, "src_ring_dev must be 4-byte aligned, n_src >= 1");
        TRY(check_image_dims(c, src_w, src_h));
        src = (const uint32_t *)src_ring_dev; W = src_w; H = src_h;
    } else {
        if (!c->img) return fail(c, HG_ERR_STATE, "no image set (hg_image_set)");
        src = c->img; W = c->W; H = c->H; n_src = 1;
    }
    const size_t T = (size_t)c->n_tris, pts_per_frame = 2 * (size_t)c->n_pts;
    const size_t slot_px = hg_stream_slot_bytes(max_out_w, max_out_h) / 4;
    const size_t bin_stride = (size_t)pwf_bins_x(max_out_w) * (size_t)max_out_h;
    TRY(ensure(c, c->dst_pts, sizeof(float) * pts_per_frame * (size_t)n_frames));
    c->map32_current = false;
    c->last_inv_len = -1;
    // chunk: frames in flight never share a ring slot, and the per-frame scratch stays below ~1.5 GB
    const size_t per_frame = (sizeof(TriRec) + 48) * T + bin_stride * (4 + 4 * PW_BIN_CAP + 32);
    int chunk = (int)((1500ull << 20) / (per_frame ? per_frame : 1));
    if (n_frames > n_slots) {
        // the stream wraps around the ring inside this call: a chunk covers at most half of it, so the deferred redo of a
        // flagged frame (one chunk later) still finds its slot untouched
        NEED(c, n_slots >= 2, "a stream longer than the ring needs n_slots >= 2");
        if (chunk > n_slots / 2) chunk = n_slots / 2;
    } else if (chunk > n_slots) {
        chunk = n_slots;
    }
    if (chunk > 1024) chunk = 1024;
    if (chunk < 1) chunk = 1;
    const bool fused = !c->force_general && c->n_tris < PW_MAX_TRIS;
    // two lanes, like hg_warp_piecewise_inverse_batch: chunk k+1 is binned beside the pixel kernel of chunk k
    const bool two = pw_two_lanes(c);
    if (fused && two && chunk > (c->pw_chunk > 0 ? c->pw_chunk : 64)) chunk = c->pw_chunk > 0 ? c->pw_chunk : 64;
    const bool lanes = fused && two && n_frames > chunk;
    for (int l = 0; l < (lanes ? 2 : 1); ++l) TRY(pw_scratch_ensure(c, pw_scratch(c, l), T, chunk, bin_stride * (size_t)chunk));
    TRY(ensure(c, c->sinfo, sizeof(StreamInfo) * (size_t)n_frames));
    TRY(ensure_pinned(c, (sizeof(StreamInfo) + sizeof(int)) * (size_t)n_frames));
    StreamInfo *h_info = (StreamInfo *)c->pin_big;
    int *h_status = (int *)((char *)c->pin_big + sizeof(StreamInfo) * (size_t)n_frames);
    for (auto &e : c->ev_chunk)
        if (!e) CU(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (lanes) TRY(pw_lanes_begin(c));
    cudaStream_t pre = lanes ? c->stream_pre : c->stream;
    // frames of chunk [f0, f0 + nf) the span bins could not express: the general, map-based path (exact for everything)
    bool any_general = false;
    auto redo_flagged = [&](int f0, int nf) -> int {
        bool first = true;
        for (int f = f0; f < f0 + nf; ++f) {
            const StreamInfo &I = h_info[f];
            if (I.status == 2) continue;  // skipped: reported to the caller
            if (!fused || h_status[f] == 1) {
                if (first && lanes) {
                    // the general path writes scratch set 0 and the map from the context stream: nothing of a later chunk
                    // may be in flight on either lane (rare: a frame the bins cannot express)
                    CU(c, cudaStreamSynchronize(c->stream_pre));
                    CU(c, cudaStreamSynchronize(c->stream));
                }
                first = false;
                PwFrameHost g{src + (n_src > 1 ? (size_t)(((long long)first_frame + f) % n_src) * (size_t)W * H : 0),
                              (uint32_t *)out_ring_dev + (size_t)I.slot * slot_px, W, H, I.x_off, I.y_off, I.o_w, I.o_h};
                TRY(pw_inverse_general_frame(c, (const float *)c->dst_pts.p + pts_per_frame * (size_t)f, g, min_src_x, min_src_y));
                c->n_general++;
                any_general = true;
            } else {
                c->n_fused++;
            }
        }
        if (!first && lanes) CU(c, cudaStreamSynchronize(c->stream));
        return HG_OK;
    };
    int prev_f0 = -1, prev_nf = 0, k = 0;
    for (int f0 = 0; f0 < n_frames; f0 += chunk, ++k) {
        const int nf = n_frames - f0 < chunk ? n_frames - f0 : chunk;
        const int lane = lanes ? (k & 1) : -1;
        const PwScratch S = pw_scratch(c, lanes ? lane : 0);
        float *dd = (float *)c->dst_pts.p + pts_per_frame * (size_t)f0;
        // this chunk's destiny points go up while the previous chunk computes (a pageable source is staged by the driver
        // before the call returns; the copy itself is stream-ordered in front of this chunk's kernels)
        CU(c, cudaMemcpyAsync(dd, dst_pts + pts_per_frame * (size_t)f0, sizeof(float) * pts_per_frame * (size_t)nf,
                              cudaMemcpyHostToDevice, pre));
        if (lanes && k >= 2) CU(c, cudaStreamWaitEvent(pre, c->ev_pix[lane], 0));  // the lane's scratch is free again
        if (!pw_band_binning(c)) CU(c, cudaMemsetAsync(S.bin_cnt->p, 0, sizeof(unsigned) * 2 * bin_stride * (size_t)nf, pre));
        CU(c, cudaMemsetAsync(S.fstatus->p, 0, sizeof(int) * (size_t)nf, pre));
        StreamArgs a{};
        a.dst_pts = dd; a.n_pts = c->n_pts; a.n_frames = nf; a.frame0 = (long long)first_frame + f0;
        a.src = src; a.src_stride_px = (size_t)W * H; a.n_src = n_src; a.W = W; a.H = H;
        a.out_ring = (uint32_t *)out_ring_dev; a.slot_px = slot_px; a.n_slots = n_slots; a.max_w = max_out_w; a.max_h = max_out_h;
        a.rec = (const TriRec *)S.rec->p; a.invd = (const float *)S.invd->p; a.yr = (const int2 *)S.yr->p;
        a.bin_cnt = (unsigned *)S.bin_cnt->p; a.bin_ent = (unsigned *)S.bin_ent->p; a.bin_run = (uint4 *)S.bin_run->p;
        a.status = (int *)S.fstatus->p; a.bin_stride = bin_stride;
        a.n_tris = c->n_tris; a.minSrcX = min_src_x; a.minSrcY = min_src_y;
        a.frames_out = (FusedFrame *)S.fframes->p;
        a.info_out = (StreamInfo *)c->sinfo.p + f0;
        pw_stream_frames_kernel<<<(unsigned)((nf + 3) / 4), 128, 0, pre>>>(a);
        c->launches++;
        CU(c, cudaGetLastError());
        if (fused) TRY(pw_fused_launch(c, S, lane, dd, nf, max_out_w, max_out_h, bin_stride, false));  // windows are decided on the device
        // the scratch is reused by a later chunk: collect this chunk's status and windows first (stream-ordered copies)
        CU(c, cudaMemcpyAsync(h_status + f0, S.fstatus->p, sizeof(int) * (size_t)nf, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaMemcpyAsync(h_info + f0, (StreamInfo *)c->sinfo.p + f0, sizeof(StreamInfo) * (size_t)nf, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaEventRecord(c->ev_chunk[k & 1], c->stream));
        if (lanes) CU(c, cudaEventRecord(c->ev_pix[lane], c->stream));
        // while this chunk runs, look at the previous one: its flagged frames are redone behind this chunk in stream
        // order, and their slots are not touched by it (a chunk covers at most half of the ring when the stream wraps)
        if (prev_f0 >= 0) {
            CU(c, cudaEventSynchronize(c->ev_chunk[(k - 1) & 1]));
            TRY(redo_flagged(prev_f0, prev_nf));
        }
        prev_f0 = f0;
        prev_nf = nf;
    }
    CU(c, cudaStreamSynchronize(c->stream));
    TRY(redo_flagged(prev_f0, prev_nf));
    if (any_general) CU(c, cudaStreamSynchronize(c->stream));
    if (info_out) memcpy(info_out, h_info, sizeof(StreamInfo) * (size_t)n_frames);
    return HG_OK;
}

/* ------------------------------------------------------------------ diagnostics */
int hg_debug_rcp_max_error(hg_ctx *c, int biased_exponent, int negative, double *max_rel_err)
{
    BIND(c);
    NEED(c, max_rel_err, "NULL argument");
    NEED(c, biased_exponent >= 1 && biased_exponent <= 2046, "biased_exponent must be a normal exponent");
    unsigned long long *d = (unsigned long long *)((char *)c->scratch.p + SC_MISC);
    CU(c, cudaMemsetAsync(d, 0, 8, c->stream));
    rcp_error_kernel<<<(1 << 20) / 256, 256, 0, c->stream>>>(biased_exponent, negative, d);
    c->launches++;
    CU(c, cudaGetLastError());
    unsigned long long bits = 0;
    CU(c, cudaMemcpyAsync(&bits, d, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    memcpy(max_rel_err, &bits, 8);
    return HG_OK;
}

int hg_debug_quotient_at_least(hg_ctx *c, const double *N, const double *D, const double *b, int n, int *out)
{
    BIND(c);
    NEED(c, N && D && b && out && n >= 1 && n <= (1 << 22), "bad argument");
    DevBuf buf;
    const size_t nb = sizeof(double) * (size_t)n;
    TRY(ensure(c, buf, 3 * nb + sizeof(int) * (size_t)n));
    char *p = (char *)buf.p;
    int st = HG_OK;
    do {
        if (cudaMemcpyAsync(p, N, nb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
            cudaMemcpyAsync(p + nb, D, nb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
            cudaMemcpyAsync(p + 2 * nb, b, nb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { st = HG_ERR_CUDA; break; }
        quotient_decision_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>((const double *)p, (const double *)(p + nb),
                                                                         (const double *)(p + 2 * nb), n, (int *)(p + 3 * nb));
        c->launches++;
        if (cudaMemcpyAsync(out, p + 3 * nb, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) st = HG_ERR_CUDA;
    } while (0);
    cudaFree(buf.p);
    if (st) return fail(c, st, "hg_debug_quotient_at_least: %s", cudaGetErrorString(cudaGetLastError()));
    return HG_OK;
}

/* ------------------------------------------------------------------ memory helpers */
int hg_dev_alloc(hg_ctx *c, size_t bytes, void **p)
{
    BIND(c);
    NEED(c, p, "dev_ptr is NULL");
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *p = nullptr;
        return fail(c, HG_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    return HG_OK;
}

int hg_dev_free(hg_ctx *c, void *p)
{
    BIND(c);
    if (p) {
        CU(c, cudaStreamSynchronize(c->stream));
        CU(c, cudaFree(p));
    }
    return HG_OK;
}

int hg_host_alloc_pinned(hg_ctx *c, size_t bytes, void **p)
{
    BIND(c);
    NEED(c, p, "host_ptr is NULL");
    cudaError_t e = cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *p = nullptr;
        return fail(c, HG_ERR_NOMEM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    return HG_OK;
}

int hg_host_alloc_pinned_ex(hg_ctx *c, size_t bytes, int write_combined, void **p)
{
    BIND(c);
    NEED(c, p, "host_ptr is NULL");
    cudaError_t e = cudaHostAlloc(p, bytes ? bytes : 1, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *p = nullptr;
        return fail(c, HG_ERR_NOMEM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    return HG_OK;
}

int hg_pcie_probe(hg_ctx *c, size_t bytes, int iters, double *h2d_gbs, double *d2h_gbs, double *bidir_gbs)
{
    BIND(c);
    NEED(c, h2d_gbs && d2h_gbs && bidir_gbs, "NULL argument");
    NEED(c, bytes >= 4096 && bytes <= (1ull << 31) && iters >= 1 && iters <= 4096, "bytes in [4 KiB, 2 GiB], iters in [1, 4096]");
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess && r != cudaSuccess) e = r; return r == cudaSuccess; };
    // buffers, streams and events are kept by the context: pinning 2 x 64 MiB takes tens of milliseconds and as long again
    // to release, so ranks that probe "at the same moment" (a barrier in front of the call) would otherwise drift apart by
    // about one copy phase and measure each other's idle link.  A first call allocates; calls after it start copying at once.
    if (c->probe_bytes < bytes) {
        hg_pcie_probe_release(c);
        ok(cudaHostAlloc(&c->probe_h[0], bytes, cudaHostAllocDefault));
        ok(cudaHostAlloc(&c->probe_h[1], bytes, cudaHostAllocDefault));
        ok(cudaMalloc(&c->probe_d[0], bytes));
        ok(cudaMalloc(&c->probe_d[1], bytes));
        for (auto &st : c->probe_s) ok(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        for (auto &ev : c->probe_e) ok(cudaEventCreate(&ev));
        if (e == cudaSuccess) {
            c->probe_bytes = bytes;
            memset(c->probe_h[0], 1, bytes);
            memset(c->probe_h[1], 2, bytes);
            ok(cudaMemsetAsync(c->probe_d[1], 3, bytes, c->probe_s[1]));
        }
    }
    if (e == cudaSuccess) {
        void *h_a = c->probe_h[0], *h_b = c->probe_h[1], *d_a = c->probe_d[0], *d_b = c->probe_d[1];
        cudaStream_t s1 = c->probe_s[0], s2 = c->probe_s[1];
        cudaEvent_t e0 = c->probe_e[0], e1 = c->probe_e[1], e2 = c->probe_e[2], e3 = c->probe_e[3];
        // warm-up of both directions
        ok(cudaMemcpyAsync(d_a, h_a, bytes, cudaMemcpyHostToDevice, s1));
        ok(cudaMemcpyAsync(h_b, d_b, bytes, cudaMemcpyDeviceToHost, s2));
        ok(cudaStreamSynchronize(s1)); ok(cudaStreamSynchronize(s2));
        float ms = 0.f;
        ok(cudaEventRecord(e0, s1));
        for (int i = 0; i < iters; ++i) ok(cudaMemcpyAsync(d_a, h_a, bytes, cudaMemcpyHostToDevice, s1));
        ok(cudaEventRecord(e1, s1));
        ok(cudaStreamSynchronize(s1));
        if (ok(cudaEventElapsedTime(&ms, e0, e1))) *h2d_gbs = (double)bytes * iters / (ms * 1e-3) / 1e9;
        ok(cudaEventRecord(e2, s2));
        for (int i = 0; i < iters; ++i) ok(cudaMemcpyAsync(h_b, d_b, bytes, cudaMemcpyDeviceToHost, s2));
        ok(cudaEventRecord(e3, s2));
        ok(cudaStreamSynchronize(s2));
        if (ok(cudaEventElapsedTime(&ms, e2, e3))) *d2h_gbs = (double)bytes * iters / (ms * 1e-3) / 1e9;
        // both directions at once: each stream timed by its own events, the pair by the longer of the two
        ok(cudaEventRecord(e0, s1));
        ok(cudaEventRecord(e2, s2));
        for (int i = 0; i < iters; ++i) {
            ok(cudaMemcpyAsync(d_a, h_a, bytes, cudaMemcpyHostToDevice, s1));
            ok(cudaMemcpyAsync(h_b, d_b, bytes, cudaMemcpyDeviceToHost, s2));
        }
        ok(cudaEventRecord(e1, s1));
        ok(cudaEventRecord(e3, s2));
        ok(cudaStreamSynchronize(s1)); ok(cudaStreamSynchronize(s2));
        float ma = 0.f, mb = 0.f;
        if (ok(cudaEventElapsedTime(&ma, e0, e1)) && ok(cudaEventElapsedTime(&mb, e2, e3)))
            *bidir_gbs = 2.0 * (double)bytes * iters / ((ma > mb ? ma : mb) * 1e-3) / 1e9;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        hg_pcie_probe_release(c);
        return fail(c, HG_ERR_CUDA, "hg_pcie_probe: %s", cudaGetErrorString(e));
    }
    return HG_OK;
}

int hg_host_free_pinned(hg_ctx *c, void *p)
{
    BIND(c);
    if (p) CU(c, cudaFreeHost(p));
    return HG_OK;
}

int hg_memcpy_h2d(hg_ctx *c, void *dst, const void *src, size_t bytes)
{
    BIND(c);
    CU(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return HG_OK;
}

int hg_memcpy_d2h(hg_ctx *c, void *dst, const void *src, size_t bytes)
{
    BIND(c);
    CU(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    return HG_OK;
}

int hg_output_device(hg_ctx *c, void **p, size_t *bytes)
{
    if (!c || !p || !bytes) return HG_ERR_INVALID;
    *p = c->out.p;
    *bytes = c->out_bytes;
    return HG_OK;
}

}  // extern "C"
