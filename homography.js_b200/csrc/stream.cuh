// stream.cuh — device-side frame bookkeeping for streamed piecewise warps, and the per-frame checksum.
//
//   pw_stream_frames_kernel   A9 for a chunk of stream frames ON THE DEVICE: output window of every frame from its
//                             destiny points (_induceBestObjectiveWidthAndHeight, H.js:706-710 + minmaxXYofArray,
//                             H.js:1558: offsets = round(min), size = round(max) - round(min)), placement of the frame
//                             in the caller's output ring, and the FusedFrame descriptor the fused piecewise kernels
//                             read — no host round trip between "new destiny points" and the pixel loop (the video
//                             protocol of test/benchmark.js:96-113: setDestinyPoints + warp per frame).
//   checksum_frames_kernel    64-bit position-weighted checksum of each frame of a batch (parity gates at sizes where
//                             the oracle can only check a sample pixel by pixel).
#pragma once
#include "piecewise_fused.cuh"

namespace hg {

// mirrors hg_stream_info (include/hgwarp.h)
struct StreamInfo {
    int x_off, y_off, o_w, o_h;
    int slot;     // ring slot the frame was written to
    int status;   // 0 = warped; 2 = skipped (empty / non-finite window, or larger than the ring slot)
};

struct StreamArgs {
    const float *dst_pts;      // chunk frames x n_pts x 2
    int n_pts, n_frames;
    long long frame0;          // stream index of the chunk's first frame (ring slot = index mod n_slots)
    const uint32_t *src;       // source image(s)
    size_t src_stride_px;      // distance between ring sources (0 = one shared image)
    int n_src;
    int W, H;
    uint32_t *out_ring;
    size_t slot_px;
    int n_slots, max_w, max_h;
    const TriRec *rec;
    const float *invd;
    const int2 *yr;
    unsigned *bin_cnt, *bin_ent;
    uint4 *bin_run;
    int *status;
    size_t bin_stride;         // bins reserved per frame
    int n_tris, minSrcX, minSrcY;
    FusedFrame *frames_out;
    StreamInfo *info_out;
};

// one warp per frame; `>` / `<` skip NaN exactly like minmaxXYofArray (H.js:1558)
__global__ void __launch_bounds__(128) pw_stream_frames_kernel(const StreamArgs a)
{
    const int f = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (f >= a.n_frames) return;
    const int lane = threadIdx.x & 31;
    const float *p = a.dst_pts + (size_t)f * 2 * a.n_pts;
    const float inf = __int_as_float(0x7f800000);
    float mnx = inf, mny = inf, mxx = -inf, mxy = -inf;
    bool nan_seen = false;
    for (int i = lane; i < a.n_pts; i += 32) {
        const float x = p[2 * i], y = p[2 * i + 1];
        nan_seen = nan_seen || (x != x) || (y != y);
        if (x > mxx) mxx = x;
        if (x < mnx) mnx = x;
        if (y > mxy) mxy = y;
        if (y < mny) mny = y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    nan_seen = __any_sync(0xffffffffu, nan_seen);
    if (lane != 0) return;
    const double x0 = js_round((double)mnx), y0 = js_round((double)mny);
    const double w = __dsub_rn(js_round((double)mxx), x0), h = __dsub_rn(js_round((double)mxy), y0);
    // a frame this stream can hold: finite window of >= 1x1 pixels inside the ring slot, offsets in the supported
    // range (the extrema then bound every point by 2^18 + 65536 < 2^20), no NaN point (the per-triangle solves
    // would propagate it; the single-frame entry points reject such meshes up front)
    const bool ok = !nan_seen && w >= 1.0 && h >= 1.0 && w <= (double)a.max_w && h <= (double)a.max_h &&
                    fabs(x0) <= 262144.0 && fabs(y0) <= 262144.0;
    const long long g = a.frame0 + f;
    const int slot = (int)(g % a.n_slots);
    FusedFrame F;
    F.src = a.src + (a.n_src > 1 ? (size_t)(g % a.n_src) * a.src_stride_px : 0);
    F.out = a.out_ring + (size_t)slot * a.slot_px;
    F.rec = a.rec + (size_t)a.n_tris * f;
    F.inv = a.invd + 8 * (size_t)a.n_tris * f;
    F.yr = a.yr + (size_t)a.n_tris * f;
    F.bin_cnt = a.bin_cnt + 2 * a.bin_stride * f;
    F.bin_ent = a.bin_ent + a.bin_stride * f * PW_BIN_CAP;
    F.bin_run = a.bin_run + 2 * a.bin_stride * f;
    F.status = a.status + f;
    F.W = a.W; F.H = a.H;
    F.xOff = ok ? (int)x0 : 0;
    F.yOff = ok ? (int)y0 : 0;
    F.oW = ok ? (int)w : 0;    // a 0 x 0 window: every kernel of the chain skips the frame
    F.oH = ok ? (int)h : 0;
    F.minSrcX = a.minSrcX; F.minSrcY = a.minSrcY;
    F.n_tris = ok ? a.n_tris : 0;
    F.bins_x = pwf_bins_x(F.oW);
    a.frames_out[f] = F;
    StreamInfo I;
    I.x_off = F.xOff; I.y_off = F.yOff; I.o_w = F.oW; I.o_h = F.oH;
    I.slot = slot;
    I.status = ok ? 0 : 2;
    a.info_out[f] = I;
    if (!ok) a.status[f] = 2;  // the status plane was zeroed before this kernel
}

// cs(frame) = sum_i  pixel_i * (((i * 2654435761) mod 2^32) | 1)  +  n * 0x9E3779B97F4A7C15      (mod 2^64)
// pixel_i = the RGBA8 pixel as a little-endian 32-bit word.  Order-independent (a sum), so any reduction tree gives
// the same value; tests/ and bench.py compute the same expression with numpy over the oracle's output.
struct ChecksumFrame {
    const uint32_t *px;
    long long n;
};

__global__ void __launch_bounds__(256) checksum_frames_kernel(const ChecksumFrame *frames, unsigned long long *out)
{
    const ChecksumFrame F = frames[blockIdx.y];
    unsigned long long acc = 0ull;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < F.n; i += stride) {
        const unsigned wgt = ((unsigned)i * 2654435761u) | 1u;
        acc += (unsigned long long)__ldg(F.px + i) * (unsigned long long)wgt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ unsigned long long s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0ull;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += s[k];
        if (blockIdx.x == 0) t += (unsigned long long)F.n * 0x9E3779B97F4A7C15ull;
        atomicAdd(out + blockIdx.y, t);
    }
}

}  // namespace hg
