/* hgwarp_napi.c — thin Node-API addon over the C ABI of include/hgwarp.h.
 *
 * Every exported function is a 1:1 marshalling stub: JS typed arrays are passed as borrowed pointers into their
 * ArrayBuffers (no copies on the JS side), results are written into a Uint8ClampedArray allocated here, and a
 * non-zero hg_status becomes a thrown JS Error carrying hg_last_error().  js/Homography.mjs holds the state machine.
 *
 * Build (with a Node toolchain):
 *   cc -O2 -fPIC -shared -DHG_USE_SYSTEM_NAPI -I$(node -p "process.execPath+'/../../include/node'") \
 *      -I../../include hgwarp_napi.c -L.. -lhgwarp -Wl,-rpath,'$ORIGIN/..' -o hgwarp.node
 * In this image (no Node) it is compiled against js/napi_min.h and executed under the miniature Node-API runtime of
 * tests/napi_mock/ (tests/test_napi_addon_mock.py).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/hgwarp.h"
#include "napi_min.h"

#define MAXARG 12
#define ARGS(n)                                                         \
    size_t argc = (n);                                                  \
    napi_value argv[MAXARG];                                            \
    if (napi_get_cb_info(env, info, &argc, argv, NULL, NULL) != napi_ok || argc < (n)) \
        return throw_text(env, "hgwarp: wrong number of arguments")

static napi_value throw_text(napi_env env, const char *msg)
{
    napi_throw_error(env, "HGWARP", msg);
    return NULL;
}

static napi_value check(napi_env env, hg_ctx *ctx, int status)
{
    if (status == HG_OK) {
        napi_value u;
        napi_get_undefined(env, &u);
        return u;
    }
    char buf[600];
    snprintf(buf, sizeof buf, "hgwarp status %d: %s", status, hg_last_error(ctx));
    return throw_text(env, buf);
}

static hg_ctx *ctx_of(napi_env env, napi_value v)
{
    void *p = NULL;
    return napi_get_value_external(env, v, &p) == napi_ok ? (hg_ctx *)p : NULL;
}

static void *typed(napi_env env, napi_value v, napi_typedarray_type want, size_t min_len)
{
    napi_typedarray_type t;
    size_t n = 0;
    void *data = NULL;
    if (napi_get_typedarray_info(env, v, &t, &n, &data, NULL, NULL) != napi_ok) return NULL;
    if (t != want && !(want == napi_uint8_array && t == napi_uint8_clamped_array)) return NULL;
    return n >= min_len ? data : NULL;
}

static int i32(napi_env env, napi_value v)
{
    int32_t x = 0;
    napi_get_value_int32(env, v, &x);
    return x;
}

static void ctx_finalize(napi_env env, void *data, void *hint)
{
    (void)env; (void)hint;
    hg_ctx_destroy((hg_ctx *)data);
}

/* createContext(device) -> external */
static napi_value CreateContext(napi_env env, napi_callback_info info)
{
    ARGS(1);
    hg_ctx *ctx = NULL;
    int st = hg_ctx_create(i32(env, argv[0]), &ctx);
    if (st != HG_OK) return check(env, NULL, st);
    napi_value ext;
    napi_create_external(env, ctx, ctx_finalize, NULL, &ext);
    return ext;
}

/* setImage(ctx, Uint8ClampedArray rgba, w, h)           <- this._image = image.data, H.js:298 */
static napi_value SetImage(napi_env env, napi_callback_info info)
{
    ARGS(4);
    hg_ctx *ctx = ctx_of(env, argv[0]);
    int w = i32(env, argv[2]), h = i32(env, argv[3]);
    const uint8_t *px = (const uint8_t *)typed(env, argv[1], napi_uint8_array, (size_t)w * h * 4);
    if (!ctx || !px) return throw_text(env, "setImage(ctx, Uint8ClampedArray, w, h)");
    return check(env, ctx, hg_image_set(ctx, px, w, h));
}

/* solveWithLimits(ctx, kind, Float64Array src, Float64Array dst, w, h) -> {matrix, limits}
 *                                                        <- calculateTransformMatrix + calculateTransformLimits */
static napi_value SolveWithLimits(napi_env env, napi_callback_info info)
{
    ARGS(6);
    hg_ctx *ctx = ctx_of(env, argv[0]);
    int kind = i32(env, argv[1]);
    size_t np = kind == HG_AFFINE ? 6 : 8;
    const double *src = (const double *)typed(env, argv[2], napi_float64_array, np);
    const double *dst = (const double *)typed(env, argv[3], napi_float64_array, np);
    double w = 0, h = 0;
    napi_get_value_double(env, argv[4], &w);
    napi_get_value_double(env, argv[5], &h);
    if (!ctx || !src || !dst) return throw_text(env, "solveWithLimits(ctx, kind, Float64Array, Float64Array, w, h)");
    napi_value ab, m, lab, lim, out;
    void *mp = NULL, *lp = NULL;
    napi_create_arraybuffer(env, kind == HG_AFFINE ? 24 : 64, &mp, &ab);
    napi_create_typedarray(env, kind == HG_AFFINE ? napi_float32_array : napi_float64_array, np, ab, 0, &m);
    napi_create_arraybuffer(env, 32, &lp, &lab);
    napi_create_typedarray(env, napi_float64_array, 4, lab, 0, &lim);
    int st = hg_solve_with_limits(ctx, kind, src, dst, w, h, mp, (double *)lp);
    if (st != HG_OK) return check(env, ctx, st);
    napi_create_object(env, &out);
    napi_set_named_property(env, out, "matrix", m);
    napi_set_named_property(env, out, "limits", lim);
    return out;
}

/* a fresh Uint8ClampedArray of ow*oh*4 bytes; NULL (with a pending exception) when the window is empty or the engine
 * cannot allocate it — the warp must not run with a NULL result buffer */
static napi_value new_output(napi_env env, int ow, int oh, uint8_t **data)
{
    napi_value ab, arr;
    void *p = NULL;
    *data = NULL;
    if (ow < 1 || oh < 1) return throw_text(env, "output window must be at least 1x1");
    if (napi_create_arraybuffer(env, (size_t)ow * oh * 4, &p, &ab) != napi_ok || !p ||
        napi_create_typedarray(env, napi_uint8_clamped_array, (size_t)ow * oh * 4, ab, 0, &arr) != napi_ok)
        return throw_text(env, "out of memory for the result image");
    *data = (uint8_t *)p;
    return arr;
}

/* the destiny points of a piecewise call: a Float32Array holding at least the 2*n_pts floats the library will read for
 * the mesh this context holds (NULL when there is no mesh or the array is too short) */
static const float *mesh_points(napi_env env, hg_ctx *ctx, napi_value v, int *n_tris)
{
    int np = 0, nt = 0;
    if (!ctx || hg_piecewise_mesh_size(ctx, &np, &nt) != HG_OK || np < 3) return NULL;
    if (n_tris) *n_tris = nt;
    return (const float *)typed(env, v, napi_float32_array, (size_t)np * 2);
}

/* warpInversePoints(ctx, kind, Float64Array dst, Float64Array src, xOff, yOff, oW, oH) -> Uint8ClampedArray
 *                                                        <- _inverseGeometricWarp, H.js:987 */
static napi_value WarpInversePoints(napi_env env, napi_callback_info info)
{
    ARGS(8);
    hg_ctx *ctx = ctx_of(env, argv[0]);
    int kind = i32(env, argv[1]);
    size_t np = kind == HG_AFFINE ? 6 : 8;
    const double *dst = (const double *)typed(env, argv[2], napi_float64_array, np);
    const double *src = (const double *)typed(env, argv[3], napi_float64_array, np);
    if (!ctx || !dst || !src) return throw_text(env, "warpInversePoints(ctx, kind, Float64Array, Float64Array, ...)");
    int ow = i32(env, argv[6]), oh = i32(env, argv[7]);
    uint8_t *out = NULL;
    napi_value arr = new_output(env, ow, oh, &out);
    if (!out) return NULL;
    int st = hg_warp_inverse_points(ctx, kind, dst, src, i32(env, argv[4]), i32(env, argv[5]), ow, oh, out, NULL);
    return st == HG_OK ? arr : check(env, ctx, st);
}

/* warpForwardMatrix(ctx, kind, matrix, xOff, yOff, oW, oH) -> Uint8ClampedArray     <- _geometricWarp, H.js:911 */
static napi_value WarpForwardMatrix(napi_env env, napi_callback_info info)
{
    ARGS(7);
    hg_ctx *ctx = ctx_of(env, argv[0]);
    int kind = i32(env, argv[1]);
    const void *m = kind == HG_AFFINE ? typed(env, argv[2], napi_float32_array, 6) : typed(env, argv[2], napi_float64_array, 8);
    if (!ctx || !m) return throw_text(env, "warpForwardMatrix(ctx, kind, matrix, ...)");
    int ow = i32(env, argv[5]), oh = i32(env, argv[6]);
    uint8_t *out = NULL;
    napi_value arr = new_output(env, ow, oh, &out);
    if (!out) return NULL;
    int st = hg_warp_forward_matrix(ctx, kind, m, i32(env, argv[3]), i32(env, argv[4]), ow, oh, out, NULL);
    return st == HG_OK ? arr : check(env, ctx, st);
}

/* setMesh(ctx, Float32Array srcPts, Uint32Array triangles)      <- this._srcPoints / this._triangles */
static napi_value SetMesh(napi_env env, napi_callback_info info)
{
    ARGS(3);
    hg_ctx *ctx = ctx_of(env, argv[0]);
    napi_typedarray_type t;
    size_t npts2 = 0, ntri3 = 0;
    void *pts = NULL, *tris = NULL;
    if (!ctx || napi_get_typedarray_info(env, argv[1], &t, &npts2, &pts, NULL, NULL) != napi_ok || t != napi_float32_array ||
        napi_get_typedarray_info(env, argv[2], &t, &ntri3, &tris, NULL, NULL) != napi_ok || t != napi_uint32_array)
        return throw_text(env, "setMesh(ctx, Float32Array, Uint32Array)");
    return check(env, ctx, hg_piecewise_set_mesh(ctx, (const float *)pts, (int)(npts2 / 2), (const uint32_t *)tris, (int)(ntri3 / 3)));
}

/* delaunay(Float32Array | Float64Array points) -> Uint32Array     <- Delaunay(points), H.js:1216 (delaunator@5.0.0) */
static napi_value Delaunay(napi_env env, napi_callback_info info)
{
    ARGS(1);
    napi_typedarray_type t;
    size_t n2 = 0;
    void *data = NULL;
    if (napi_get_typedarray_info(env, argv[0], &t, &n2, &data, NULL, NULL) != napi_ok ||
        (t != napi_float32_array && t != napi_float64_array))
        return throw_text(env, "delaunay(Float32Array | Float64Array)");
    const int n = (int)(n2 / 2);
    /* the package reads coords[i] as JS Numbers: widen a Float32Array exactly */
    double *pts = (double *)malloc(sizeof(double) * (n2 ? n2 : 1));
    if (!pts) return throw_text(env, "out of memory");
    for (size_t i = 0; i < n2; ++i) pts[i] = t == napi_float32_array ? (double)((const float *)data)[i] : ((const double *)data)[i];
    const int cap = 2 * n > 5 ? 2 * n - 5 : 0;
    napi_value ab, arr;
    void *p = NULL;
    if (napi_create_arraybuffer(env, (size_t)cap * 12, &p, &ab) != napi_ok || (cap > 0 && !p)) {
        free(pts);
        return throw_text(env, "out of memory");
    }
    int nt = 0;
    const int st = hg_delaunay(pts, n, (uint32_t *)p, cap, &nt);
    free(pts);
    if (st != HG_OK) return throw_text(env, "hg_delaunay failed");
    /* a view of the first nt triangles, like the package's `triangles` subarray */
    napi_create_typedarray(env, napi_uint32_array, (size_t)nt * 3, ab, 0, &arr);
    return arr;
}

/* pngDecode(Uint8Array file) -> {data: Uint8ClampedArray, width, height}     <- canvas drawImage + getImageData,
 *                                                                               H.js:1071-1076 / loadImage in nodeTest.js */
static napi_value PngDecode(napi_env env, napi_callback_info info)
{
    ARGS(1);
    napi_typedarray_type t;
    size_t n = 0;
    void *data = NULL;
    if (napi_get_typedarray_info(env, argv[0], &t, &n, &data, NULL, NULL) != napi_ok || (t != napi_uint8_array && t != napi_uint8_clamped_array))
        return throw_text(env, "pngDecode(Uint8Array)");
    int w = 0, h = 0;
    if (hg_png_decode((const uint8_t *)data, n, NULL, 0, &w, &h) != HG_OK) return throw_text(env, "pngDecode: not a PNG this decoder supports");
    uint8_t *px = NULL;
    napi_value arr = new_output(env, w, h, &px), obj, vw, vh;
    if (!px) return NULL;
    if (hg_png_decode((const uint8_t *)data, n, px, (size_t)w * h * 4, &w, &h) != HG_OK) return throw_text(env, "pngDecode: malformed image data");
    napi_create_object(env, &obj);
    napi_create_int32(env, w, &vw);
    napi_create_int32(env, h, &vh);
    napi_set_named_property(env, obj, "data", arr);
    napi_set_named_property(env, obj, "width", vw);
    napi_set_named_property(env, obj, "height", vh);
    return obj;
}

/* jpegDecode(Uint8Array file) -> {data: Uint8ClampedArray, width, height}    <- the same canvas path for a JPEG file */
static napi_value JpegDecode(napi_env env, napi_callback_info info)
{
    ARGS(1);
    napi_typedarray_type t;
    size_t n = 0;
    void *data = NULL;
    if (napi_get_typedarray_info(env, argv[0], &t, &n, &data, NULL, NULL) != napi_ok || (t != napi_uint8_array && t != napi_uint8_clamped_array))
        return throw_text(env, "jpegDecode(Uint8Array)");
    int w = 0, h = 0;
    int st = hg_jpeg_decode((const uint8_t *)data, n, NULL, 0, &w, &h);
    if (st != HG_OK)
        return throw_text(env, st == HG_ERR_UNSUPPORTED ? "jpegDecode: a JPEG mode this decoder does not support (CMYK, arithmetic coding, ...)"
                                                        : "jpegDecode: not a JPEG file");
    uint8_t *px = NULL;
    napi_value arr = new_output(env, w, h, &px), obj, vw, vh;
    if (!px) return NULL;
    st = hg_jpeg_decode((const uint8_t *)data, n, px, (size_t)w * h * 4, &w, &h);
    if (st != HG_OK)
        return throw_text(env, st == HG_ERR_UNSUPPORTED ? "jpegDecode: a JPEG mode this decoder does not support (CMYK, arithmetic coding, ...)"
                                                        : "jpegDecode: malformed image data");
    napi_create_object(env, &obj);
    napi_create_int32(env, w, &vw);
    napi_create_int32(env, h, &vh);
    napi_set_named_property(env, obj, "data", arr);
    napi_set_named_property(env, obj, "width", vw);
    napi_set_named_property(env, obj, "height", vh);
    return obj;
}

/* pngEncode(Uint8ClampedArray rgba, width, height) -> Uint8Array file        <- toDataURL, H.js:480-483 */
static napi_value PngEncode(napi_env env, napi_callback_info info)
{
    ARGS(3);
    const int w = i32(env, argv[1]), h = i32(env, argv[2]);
    const uint8_t *rgba = (w > 0 && h > 0) ? (const uint8_t *)typed(env, argv[0], napi_uint8_array, (size_t)w * h * 4) : NULL;
    if (!rgba) return throw_text(env, "pngEncode(Uint8ClampedArray, width, height)");
    const size_t cap = hg_png_encode_bound(w, h);
    uint8_t *tmp = (uint8_t *)malloc(cap ? cap : 1);
    size_t n = 0;
    if (!tmp || hg_png_encode(rgba, w, h, tmp, cap, &n) != HG_OK) {
        free(tmp);
        return throw_text(env, "pngEncode failed");
    }
    napi_value ab, arr;
    void *p = NULL;
    napi_create_arraybuffer(env, n, &p, &ab);
    memcpy(p, tmp, n);
    free(tmp);
    napi_create_typedarray(env, napi_uint8_array, n, ab, 0, &arr);
    return arr;
}

/* piecewiseMatrices(ctx, Float32Array dstPts, nTris) -> Float32Array(6*nTris)
 *                                                        <- _calculatePiecewiseAffineTransformMatrices, H.js:785 */
static napi_value PiecewiseMatrices(napi_env env, napi_callback_info info)
{
    ARGS(3);
    hg_ctx *ctx = ctx_of(env, argv[0]);
    int mesh_tris = -1;
    const float *dst = mesh_points(env, ctx, argv[1], &mesh_tris);
    int nt = i32(env, argv[2]);
    /* the library writes 6 floats per triangle OF THE CONTEXT MESH: the caller's count must be that one */
    if (!ctx || !dst || nt < 0 || nt != mesh_tris)
        return throw_text(env, "piecewiseMatrices(ctx, Float32Array dstPts (2 per mesh point), nTris of the mesh)");
    napi_value ab, arr;
    void *p = NULL;
    if (napi_create_arraybuffer(env, (size_t)nt * 24, &p, &ab) != napi_ok || (nt > 0 && !p) ||
        napi_create_typedarray(env, napi_float32_array, (size_t)nt * 6, ab, 0, &arr) != napi_ok)
        return throw_text(env, "out of memory");
    int st = hg_piecewise_matrices(ctx, dst, (float *)p, NULL);
    return st == HG_OK ? arr : check(env, ctx, st);
}

/* warpPiecewiseInverse(ctx, Float32Array dst, xOff, yOff, oW, oH, minSrcX, minSrcY) -> Uint8ClampedArray
 *                                                        <- _inversePiecewiseAffineWarp, H.js:1029 */
static napi_value WarpPiecewiseInverse(napi_env env, napi_callback_info info)
{
    ARGS(8);
    hg_ctx *ctx = ctx_of(env, argv[0]);
    const float *dst = mesh_points(env, ctx, argv[1], NULL);
    if (!ctx || !dst) return throw_text(env, "warpPiecewiseInverse(ctx, Float32Array dstPts (2 per mesh point), ...)");
    int ow = i32(env, argv[4]), oh = i32(env, argv[5]);
    uint8_t *out = NULL;
    napi_value arr = new_output(env, ow, oh, &out);
    if (!out) return NULL;
    int st = hg_warp_piecewise_inverse(ctx, dst, i32(env, argv[2]), i32(env, argv[3]), ow, oh, i32(env, argv[6]),
                                       i32(env, argv[7]), out, NULL);
    return st == HG_OK ? arr : check(env, ctx, st);
}

/* warpPiecewiseForward(ctx, Float32Array dst, xOff, yOff, oW, oH, minSrcX, minSrcY, maxSrcX, maxSrcY, useInverseMap)
 *                                                        <- _piecewiseAffineWarp, H.js:948 */
static napi_value WarpPiecewiseForward(napi_env env, napi_callback_info info)
{
    ARGS(11);
    hg_ctx *ctx = ctx_of(env, argv[0]);
    const float *dst = mesh_points(env, ctx, argv[1], NULL);
    if (!ctx || !dst) return throw_text(env, "warpPiecewiseForward(ctx, Float32Array dstPts (2 per mesh point), ...)");
    int ow = i32(env, argv[4]), oh = i32(env, argv[5]);
    uint8_t *out = NULL;
    napi_value arr = new_output(env, ow, oh, &out);
    if (!out) return NULL;
    int st = hg_warp_piecewise_forward(ctx, dst, i32(env, argv[2]), i32(env, argv[3]), ow, oh, i32(env, argv[6]),
                                       i32(env, argv[7]), i32(env, argv[8]), i32(env, argv[9]), i32(env, argv[10]), out, NULL);
    return st == HG_OK ? arr : check(env, ctx, st);
}

static napi_value Init(napi_env env, napi_value exports)
{
    const napi_property_descriptor d[] = {
        {"createContext", 0, CreateContext, 0, 0, 0, napi_default, 0},
        {"setImage", 0, SetImage, 0, 0, 0, napi_default, 0},
        {"solveWithLimits", 0, SolveWithLimits, 0, 0, 0, napi_default, 0},
        {"warpInversePoints", 0, WarpInversePoints, 0, 0, 0, napi_default, 0},
        {"warpForwardMatrix", 0, WarpForwardMatrix, 0, 0, 0, napi_default, 0},
        {"setMesh", 0, SetMesh, 0, 0, 0, napi_default, 0},
        {"delaunay", 0, Delaunay, 0, 0, 0, napi_default, 0},
        {"pngDecode", 0, PngDecode, 0, 0, 0, napi_default, 0},
        {"jpegDecode", 0, JpegDecode, 0, 0, 0, napi_default, 0},
        {"pngEncode", 0, PngEncode, 0, 0, 0, napi_default, 0},
        {"piecewiseMatrices", 0, PiecewiseMatrices, 0, 0, 0, napi_default, 0},
        {"warpPiecewiseInverse", 0, WarpPiecewiseInverse, 0, 0, 0, napi_default, 0},
        {"warpPiecewiseForward", 0, WarpPiecewiseForward, 0, 0, 0, napi_default, 0},
    };
    napi_define_properties(env, exports, sizeof d / sizeof d[0], d);
    return exports;
}

NAPI_MODULE(hgwarp, Init)
