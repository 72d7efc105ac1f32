/*
 * hgwarp.h — C ABI of libhgwarp.so, the sm_100a image-warp engine behind the
 * Homography.js class surface.
 *
 * The reference (Eric-Canas/Homography.js, one ES module, cited as H.js:<line>) has no FFI:
 * its replaceable seam is the set of private methods / free functions that `warp()` (H.js:408)
 * and the setters call once the state is prepared.  Each entry point below names the reference
 * function it replaces.  Signatures are plain C (pointers + sizes, no CUDA / torch types) so an
 * N-API addon, cgo, JNI or ctypes can bind them directly.
 *
 * Conventions
 *   - every function returns an hg_status (0 = ok); hg_last_error(ctx) gives the text;
 *   - inputs are borrowed for the duration of the call, nothing host-side is retained;
 *   - images are row-major RGBA8 (the layout of ImageData.data / Uint8ClampedArray);
 *   - "points" are [x0,y0,x1,y1,...]; affine / projective solves take doubles (the reference
 *     accepts Float32Array or Float64Array, H.js:220), piecewise takes floats (the reference
 *     copies every triangle through a Float32Array(6) scratch, H.js:121-122,791-800);
 *   - `out_host` (may be NULL) receives oW*oH*4 bytes after a stream sync; `out_dev` (may be
 *     NULL) is a caller-owned, 16-byte-aligned device buffer written instead of the context's
 *     own output buffer.  With both NULL the result stays device-resident in the context;
 *   - a context is bound to one GPU and one CUDA stream and is not thread-safe; use one per GPU;
 *   - degenerate transforms (NaN / Inf matrices) are NOT errors: like the reference they give an
 *     all-transparent image.
 * Environment (read once per hg_ctx_create; none is needed in production):
 *   HG_GEO_STAGED=1     affine / projective inverse warps run the TMA-staged kernel (slower than the default direct-gather
 *                       kernel on every workload measured); HG_GEO_BOX_BYTES / HG_GEO_STAGES / HG_GEO_CTAS /
 *                       HG_GEO_NITER size its shared-memory ring, HG_GEO_VERBOSE=1 prints the chosen configuration,
 *                       HG_GEO_DEBUG=1 traces every ring entry.
 *   HG_GEO_NO_TALL=1    keep the default thread layout for quarter-turn maps (A/B runs).
 *   HG_BILINEAR_V1=1    HG_BILINEAR warps run the first-generation per-pixel kernel (A/B runs).
 * Supported ranges (anything else returns HG_ERR_UNSUPPORTED, never a wrong image):
 *   1 <= W,H,oW,oH <= 65536, W*H and oW*oH < 2^31, |xOff|,|yOff|,|minSrc*| <= 2^18.
 */
#ifndef HGWARP_H
#define HGWARP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define HGWARP_ABI_VERSION 1

typedef struct hg_ctx hg_ctx;

typedef enum {
    HG_OK = 0,
    HG_ERR_INVALID = 1,     /* bad argument */
    HG_ERR_CUDA = 2,        /* CUDA runtime failure (text in hg_last_error) */
    HG_ERR_NOMEM = 3,
    HG_ERR_UNSUPPORTED = 4, /* outside the supported ranges above */
    HG_ERR_STATE = 5        /* e.g. warp before hg_image_set */
} hg_status;

typedef enum { HG_AFFINE = 0, HG_PROJECTIVE = 1 } hg_kind;

/* HG_NEAREST is the reference's sampling (Math.round, H.js:1005).  HG_BILINEAR is an EXTENSION the reference does not
 * have (north_star asks for it): it applies to hg_warp_inverse_matrix / _points / _batch only, is defined by the oracle
 * (oracle/hg_oracle.c) and is accurate to <= 1 LSB per channel, not bit-exact. */
typedef enum { HG_NEAREST = 0, HG_BILINEAR = 1 } hg_sampling;

/* one frame of a batched warp: where it reads, where it writes, its output window */
typedef struct {
    const void *src_dev;  /* RGBA8 source image on the device (NULL = the context image) */
    void *out_dev;        /* 16-byte aligned device buffer of o_w*o_h*4 bytes */
    int32_t src_w, src_h; /* ignored when src_dev is NULL */
    int32_t x_off, y_off, o_w, o_h;
} hg_frame;

/* ------------------------------------------------------------------ context */
int hg_abi_version(void);
int hg_device_count(int *count);
int hg_ctx_create(int device, hg_ctx **out);
int hg_ctx_destroy(hg_ctx *ctx);
const char *hg_last_error(hg_ctx *ctx); /* ctx may be NULL: error of the last failed create */
int hg_ctx_synchronize(hg_ctx *ctx);
int hg_ctx_stream(hg_ctx *ctx, void **cuda_stream); /* the cudaStream_t all work is enqueued on */
/* CUDA-event stopwatch on the context stream (device time, not wall clock) */
int hg_timer_start(hg_ctx *ctx);
int hg_timer_stop(hg_ctx *ctx, float *elapsed_ms);
int hg_ctx_set_sampling(hg_ctx *ctx, int sampling); /* HG_NEAREST (default) | HG_BILINEAR */
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int hg_launch_count(hg_ctx *ctx, uint64_t *count);
/* per-kernel device timing of the pixel-loop kernels: enable, run, then read the summed CUDA-event
 * duration and the number of kernels it covers (resets the accumulation) */
int hg_profile_enable(hg_ctx *ctx, int on);
int hg_profile_read(hg_ctx *ctx, double *total_ms, uint64_t *n_kernels);

/* ------------------------------------------------------------------ image (this._image, H.js:298) */
/* H2D copy, stays resident.  Pageable memory is staged before the call returns; a PINNED buffer is copied asynchronously on
 * the context stream and must stay unchanged until the next call that synchronizes (a warp with out_host, hg_ctx_synchronize) */
int hg_image_set(hg_ctx *ctx, const uint8_t *rgba_host, int w, int h);
int hg_image_set_device(hg_ctx *ctx, const void *rgba_dev, int w, int h);   /* borrow a device buffer */

/* ------------------------------------------------------------------ transform solves (one in-register kernel) */
/* affineMatrixFromTriangles, H.js:1265 (f64 math, result rounded to Float32Array(6)) */
int hg_solve_affine(hg_ctx *ctx, const double src[6], const double dst[6], float out[6]);
/* projectiveMatrixFromSquares + numeric.js LU/LUsolve, H.js:1320 + 1650-1751 (Array(8) of f64) */
int hg_solve_projective(hg_ctx *ctx, const double src[8], const double dst[8], double out[8]);
/* inverseAffineMatrix, H.js:1345 */
int hg_inverse_affine(hg_ctx *ctx, const float m[6], float out[6]);
/* calculateTransformLimits, H.js:1503: out = [xOff, yOff, oW, oH] as JS Numbers (may be NaN) */
int hg_transform_limits(hg_ctx *ctx, int kind, const void *matrix, double w, double h, double out[4]);
/* calculateTransformMatrix(kind, src, dst) + calculateTransformLimits in one submission
 * (what setDestinyPoints needs, H.js:357 + 365); matrix_out: float[6] or double[8] */
int hg_solve_with_limits(hg_ctx *ctx, int kind, const double *src, const double *dst, double w, double h,
                         void *matrix_out, double limits_out[4]);

/* ------------------------------------------------------------------ affine / projective warps */
/* pixel loop of _inverseGeometricWarp, H.js:997-1011, with a given inverse (dst->src) matrix
 * (float[6] for HG_AFFINE, double[8] for HG_PROJECTIVE) */
int hg_warp_inverse_matrix(hg_ctx *ctx, int kind, const void *inv_matrix, int x_off, int y_off, int o_w, int o_h,
                           uint8_t *out_host, void *out_dev);
/* the whole of _inverseGeometricWarp, H.js:987-1013: solve calculateTransformMatrix(kind, dst, src)
 * on the device, then the pixel loop, with no host round trip in between */
int hg_warp_inverse_points(hg_ctx *ctx, int kind, const double *dst_pts, const double *src_pts, int x_off,
                           int y_off, int o_w, int o_h, uint8_t *out_host, void *out_dev);
/* _geometricWarp, H.js:911-932 (forward scatter, source raster order, last writer wins),
 * matrix = forward (src->dst) matrix */
int hg_warp_forward_matrix(hg_ctx *ctx, int kind, const void *fwd_matrix, int x_off, int y_off, int o_w, int o_h,
                           uint8_t *out_host, void *out_dev);

/* ------------------------------------------------------------------ triangulation (host only, no context, no GPU) */
/* Delaunay(points) = new Delaunator(points).triangles, H.js:1216-1218.  The package (delaunator@5.0.0, imported at
 * H.js:27, pinned by package.json:10-12 / package-lock.json:17-24) is third-party and absent from the reference tree;
 * csrc/delaunay_host.cuh restates its published algorithm so that triangle order and vertex order follow it ("parity
 * unpinned": no reference fixture holds a triangulation).  points: n_points (x, y) pairs, read as the JS Numbers a
 * Float32Array / Float64Array element converts to.  triangles_out receives 3 vertex ids per triangle, at most
 * max(2 n - 5, 0) triangles; *n_triangles is the count (0 for fewer than 3 points or collinear input).  Returns
 * HG_ERR_INVALID when capacity_triangles is too small (nothing is written). */
int hg_delaunay(const double *points, int n_points, uint32_t *triangles_out, int capacity_triangles, int *n_triangles);

/* ------------------------------------------------------------------ image files (host only, no context, no GPU) */
/* The step either side of the path when there is no canvas: the reference reads its pixels through
 * drawImage + getImageData (H.js:1071-1076) and returns results as a PNG data URL (toDataURL, H.js:480-483); its Node
 * smoke test loads test/testImgLogoBlack.png (test/nodeTest.js:11).  PNG -> the RGBA8 layout of ImageData.data: every
 * colour type and bit depth, Adam7 interlacing included (16-bit samples keep their high byte, palette / tRNS applied).
 * rgba_out == NULL: only *w / *h are filled.  HG_ERR_INVALID: malformed or unsupported file, or capacity
 * smaller than w*h*4. */
int hg_png_decode(const uint8_t *png, size_t png_bytes, uint8_t *rgba_out, size_t capacity_bytes, int *w, int *h);
/* JPEG -> RGBA8 (alpha 255), the bytes getImageData returns for the file: sequential and progressive Huffman JPEG, 8-bit,
 * grey or three components (YCbCr / RGB), any scan layout, restart intervals; accurate integer IDCT, libjpeg's default ("fancy")
 * chroma upsampling and fixed-point colour conversion (csrc/jpeg_host.cuh; checked byte for byte against libjpeg-turbo
 * through Pillow).  The Exif Orientation tag is applied the way a browser draws the image (w / h are those of the picture as
 * shown).  Same calling convention as hg_png_decode.  HG_ERR_UNSUPPORTED: lossless / arithmetic-coded /
 * 12-bit / four-component files, 1:2 vertical-only subsampling, more than 2^28 pixels; HG_ERR_INVALID: malformed data or
 * capacity too small.  The bytes may be hostile: decoding work is bounded by the file size (at most 100 scans, at most 256
 * block visits per file byte), so a small crafted file cannot buy minutes of CPU or gigabytes of memory.  Files whose scans
 * have not all arrived (truncated progressive files) decode without libjpeg's inter-block smoothing: byte-exactness with a
 * browser holds for complete files. */
int hg_jpeg_decode(const uint8_t *jpg, size_t jpg_bytes, uint8_t *rgba_out, size_t capacity_bytes, int *w, int *h);
/* RGBA8 -> PNG (8-bit RGBA, zlib level 6).  hg_png_encode_bound gives a capacity that always suffices. */
size_t hg_png_encode_bound(int w, int h);
int hg_png_encode(const uint8_t *rgba, int w, int h, uint8_t *png_out, size_t capacity_bytes, size_t *png_bytes);

/* ------------------------------------------------------------------ piecewise affine */
/* mesh = this._srcPoints (pixel range) + this._triangles (H.js:1216 / setTriangles H.js:517) */
int hg_piecewise_set_mesh(hg_ctx *ctx, const float *src_pts, int n_pts, const uint32_t *tris, int n_tris);
/* number of points / triangles of the mesh the context holds (0, 0 before hg_piecewise_set_mesh): what bindings check the
 * lengths of caller-owned point arrays and result buffers against */
int hg_piecewise_mesh_size(hg_ctx *ctx, int *n_pts, int *n_tris);
/* _calculatePiecewiseAffineTransformMatrices, H.js:785: T forward 2x3 float matrices; optionally
 * also their inverses (inverseAffineMatrix, H.js:1036-1038).  Either output may be NULL. */
int hg_piecewise_matrices(hg_ctx *ctx, const float *dst_pts, float *fwd_out, float *inv_out);
/* output window of piecewise frames with pixel-range destiny points (_induceBestObjectiveWidthAndHeight, H.js:706-710 +
 * minmaxXYofArray, H.js:1558): out[4*f..] = {xOff, yOff, oW, oH} as JS Numbers; dst_pts = n_frames x n_pts x 2 floats */
int hg_piecewise_extents(hg_ctx *ctx, const float *dst_pts, int n_pts, int n_frames, double *out);
/* _build(Inverse)TrianglesCorrespondencesMatrix + fillTriangle, H.js:817/845/1111: Int16 map of
 * map_len entries built from `pts` (n_pts of the mesh) with row stride map_width and row origin
 * y_offset; copied to map_out_host */
int hg_build_index_map(hg_ctx *ctx, const float *pts, double map_width, double y_offset, int64_t map_len,
                       int16_t *map_out_host);
/* the whole of _inversePiecewiseAffineWarp, H.js:1029-1058: per-triangle forward+inverse matrices,
 * the inverse index map and the pixel loop, all on the device */
int hg_warp_piecewise_inverse(hg_ctx *ctx, const float *dst_pts, int x_off, int y_off, int o_w, int o_h,
                              int min_src_x, int min_src_y, uint8_t *out_host, void *out_dev);
/* _piecewiseAffineWarp, H.js:948-972 (forward scatter over the source-point bounding box).
 * use_inverse_map != 0 reproduces the reference's map aliasing (the forward loop reading the map
 * left by the last inverse warp, H.js:759/847/957): the map is then the one built by the last
 * hg_warp_piecewise_inverse / hg_build_index_map call on this context. */
int hg_warp_piecewise_forward(hg_ctx *ctx, const float *dst_pts, int x_off, int y_off, int o_w, int o_h,
                              int min_src_x, int min_src_y, int max_src_x, int max_src_y, int use_inverse_map,
                              uint8_t *out_host, void *out_dev);

/* ------------------------------------------------------------------ batched / streamed frames */
/* n_frames independent inverse warps in ONE launch (grid.y = frame).  matrices: n_frames x
 * (float[6] | double[8]) on the HOST; frames on the HOST.  Results stay on the device. */
int hg_warp_inverse_batch(hg_ctx *ctx, int kind, const void *inv_matrices, const hg_frame *frames, int n_frames);
/* n_frames inverse piecewise warps sharing the context mesh, frame f using dst_pts + f*2*n_pts
 * (HOST floats); per-frame extents come from frames[f]; min_src_* as in the single call. */
int hg_warp_piecewise_inverse_batch(hg_ctx *ctx, const float *dst_pts, const hg_frame *frames, int n_frames,
                                    int min_src_x, int min_src_y);

/* n_frames complete _inverseGeometricWarp calls (H.js:987-1013) in TWO launches: frame f solves
 * calculateTransformMatrix(kind, dst_pts_f, src_pts_f) on the device (H.js:994; points are n_frames x (6 | 8) HOST
 * doubles) and the pixel loop of all frames follows with no host round trip — the per-frame setDestinyPoints() + warp()
 * protocol of test/benchmark.js:96-113 for a batch of independent frames.  Results stay on the device. */
int hg_warp_inverse_points_batch(hg_ctx *ctx, int kind, const double *dst_pts, const double *src_pts,
                                 const hg_frame *frames, int n_frames);
/* _geometricWarp (H.js:911-932) for n_frames independent frames: fwd_matrices = n_frames x (float[6] | double[8]) on
 * the HOST, frames[f].src_dev / out_dev / window as in hg_warp_inverse_batch.  Results stay on the device. */
int hg_warp_forward_batch(hg_ctx *ctx, int kind, const void *fwd_matrices, const hg_frame *frames, int n_frames);
/* _piecewiseAffineWarp (H.js:948-972) for n_frames frames sharing the context mesh and its forward map (built once
 * per source-point set, H.js:759/817), frame f using dst_pts + f*2*n_pts; min/max_src_* as in the single call. */
int hg_warp_piecewise_forward_batch(hg_ctx *ctx, const float *dst_pts, const hg_frame *frames, int n_frames,
                                    int min_src_x, int min_src_y, int max_src_x, int max_src_y);

/* ------------------------------------------------------------------ streamed piecewise frames (video) */
/* per-frame result of hg_warp_piecewise_stream */
typedef struct {
    int32_t x_off, y_off, o_w, o_h; /* the frame's output window (H.js:706-710), computed on the device */
    int32_t slot;                   /* ring slot holding its o_w*o_h*4 bytes (dense rows of o_w pixels) */
    int32_t status;                 /* 0 = warped; 2 = skipped: empty / non-finite window, or larger than a ring slot */
} hg_stream_info;
/* The video protocol of the reference (README "250 different transforms", test/benchmark.js:96-113): one mesh, a new set
 * of destiny points per frame, setDestinyPoints() + warp() per frame.  For frames first_frame .. first_frame+n_frames-1 of
 * a stream, everything a frame needs happens on the device, in order: its output window from its destiny points
 * (_induceBestObjectiveWidthAndHeight, H.js:706-710), its placement in the caller's output ring (slot = frame index mod
 * n_slots, each slot max_out_w*max_out_h*4 bytes rounded up to 256), the per-triangle matrices and their inverses, the
 * inverse triangle map and the pixel loop (_inversePiecewiseAffineWarp, H.js:1029-1058).  dst_pts: n_frames x n_pts x 2
 * HOST floats.  src_ring_dev == NULL: every frame reads the context image; otherwise frame i reads image (i mod n_src)
 * of n_src W x H images stored back to back.  info_out (HOST, n_frames entries, may be NULL) receives each frame's window
 * and slot.  Returns after the stream has been synchronized (frames a compact span list cannot express are redone by the
 * general map-based path before the call returns, never approximated). */
int hg_warp_piecewise_stream(hg_ctx *ctx, const float *dst_pts, int n_frames, int64_t first_frame, int min_src_x,
                             int min_src_y, const void *src_ring_dev, int n_src, int src_w, int src_h, void *out_ring_dev,
                             int n_slots, int max_out_w, int max_out_h, hg_stream_info *info_out);
/* bytes of one ring slot for hg_warp_piecewise_stream */
size_t hg_stream_slot_bytes(int max_out_w, int max_out_h);

/* 64-bit checksum of each frame's output (frames[f].out_dev, o_w*o_h pixels):
 *   sum_i pixel_i * (((i * 2654435761) mod 2^32) | 1) + n * 0x9E3779B97F4A7C15  (mod 2^64), pixel_i little-endian RGBA8.
 * For parity gates over batches too large to compare pixel by pixel.  out_host: n_frames values. */
int hg_checksum_frames(hg_ctx *ctx, const hg_frame *frames, int n_frames, uint64_t *out_host);

/* ------------------------------------------------------------------ pipelined host-to-host stream (video) */
/* Independent frames arriving in HOST memory (pinned for full overlap) and leaving to HOST memory: up to `depth`
 * frames in flight, H2D of frame i+1, solve+warp of frame i and D2H of frame i-1 overlap on three CUDA streams.
 * Each submit is _inverseGeometricWarp (H.js:987) for its own image and its own point sets. */
typedef struct hg_pipe hg_pipe;
int hg_pipe_create(hg_ctx *ctx, int kind, int src_w, int src_h, int max_out_w, int max_out_h, int depth, hg_pipe **out);
int hg_pipe_submit(hg_pipe *pipe, const uint8_t *rgba_host, const double *dst_pts, const double *src_pts, int x_off,
                   int y_off, int o_w, int o_h, uint8_t *out_host, uint64_t *ticket);
int hg_pipe_wait(hg_pipe *pipe, uint64_t ticket); /* returns once that frame's out_host is complete */
int hg_pipe_flush(hg_pipe *pipe);                 /* returns once every submitted frame is complete */
int hg_pipe_destroy(hg_pipe *pipe);
/* The same pipeline for piecewise frames (hg_warp_piecewise_inverse per frame, mesh = the context mesh at submit time):
 * rgba_host == NULL warps the context image (hg_image_set) — the reference's video protocol keeps one image and moves
 * the destiny points.  The output window is computed on the host exactly as H.js:706-710 does and
 * returned in window_out[4] = {x_off, y_off, o_w, o_h}; out_host must hold max_out_w*max_out_h*4 bytes.  Returns
 * HG_ERR_UNSUPPORTED (nothing submitted) when the window is empty or exceeds the pipe's maximum. */
int hg_pipe_create_piecewise(hg_ctx *ctx, int src_w, int src_h, int max_out_w, int max_out_h, int depth, hg_pipe **out);
int hg_pipe_submit_piecewise(hg_pipe *pipe, const uint8_t *rgba_host, const float *dst_pts, int min_src_x, int min_src_y,
                             uint8_t *out_host, int32_t window_out[4], uint64_t *ticket);

/* ------------------------------------------------------------------ diagnostics */
/* max over all 2^20 high-mantissa patterns (x 6 low words) of |1 - d*r|, r = the reciprocal the projective
 * fast path uses (MUFU.RCP64H + one Newton step), for doubles with the given biased exponent / sign */
int hg_debug_rcp_max_error(hg_ctx *ctx, int biased_exponent, int negative, double *max_rel_err);
/* out[i] = 1 iff the device's division-free decision says RN(N[i] / D[i]) >= b[i] (b a non-zero multiple of 1/2,
 * |b| < 2^19; D finite, non-zero): the exact resolution of projective pixels on a decision boundary (warp_geo.cuh) */
int hg_debug_quotient_at_least(hg_ctx *ctx, const double *N, const double *D, const double *b, int n, int *out);
/* on != 0: inverse piecewise warps always take the general map-based path (tests compare it with the fused one) */
int hg_debug_force_general(hg_ctx *ctx, int on);
/* how the fused path bins a frame's triangle rows: 0 = automatic, 1 = span + run passes over global bins, 2 = one pass per
 * band of map rows through shared memory (tests run every frame through both and compare) */
int hg_debug_piecewise_binning(hg_ctx *ctx, int mode);
/* how many inverse piecewise frames were finished by the fused (map-free) path / by the general map-based path */
int hg_debug_piecewise_stats(hg_ctx *ctx, uint64_t *frames_fused, uint64_t *frames_general);

/* ------------------------------------------------------------------ device memory helpers (benchmarks / bindings) */
int hg_dev_alloc(hg_ctx *ctx, size_t bytes, void **dev_ptr);
int hg_dev_free(hg_ctx *ctx, void *dev_ptr);
int hg_host_alloc_pinned(hg_ctx *ctx, size_t bytes, void **host_ptr);
int hg_host_free_pinned(hg_ctx *ctx, void *host_ptr);
/* write_combined != 0: cudaHostAllocWriteCombined — for buffers the host only writes and the GPU's copy engine only
 * reads (a source ring): reads over PCIe skip the CPU cache snoop */
int hg_host_alloc_pinned_ex(hg_ctx *ctx, size_t bytes, int write_combined, void **host_ptr);
/* Raw copy-engine ceiling of this GPU's host link, for judging end-to-end numbers: `iters` pinned copies of `bytes`
 * host->device alone, device->host alone, and both directions at once on two streams; GB/s each (bidir = sum of both
 * directions).  Several processes calling this at the same moment (one per GPU) measure the box's shared host fabric. */
int hg_pcie_probe(hg_ctx *ctx, size_t bytes, int iters, double *h2d_gbs, double *d2h_gbs, double *bidir_gbs);
int hg_memcpy_h2d(hg_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes); /* async on the ctx stream */
int hg_memcpy_d2h(hg_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes); /* async on the ctx stream */
/* the context's own output buffer of the last non-batched warp (device pointer + byte size) */
int hg_output_device(hg_ctx *ctx, void **dev_ptr, size_t *bytes);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* HGWARP_H */
