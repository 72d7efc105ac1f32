/* hgwarp_cpu_double.c — TEST DOUBLE of the GPU entry points the Node-API addon binds (include/hgwarp.h), answered by the
 * CPU oracle (oracle/libhgoracle.so).  It exists so that the chain
 *     class surface -> native.* -> N-API marshalling (js/hgwarp_napi.c) -> C ABI
 * can be executed and checked on a machine WITHOUT a GPU: every argument the addon forwards (order, type, units) lands in
 * an oracle call whose result the tests compare with the reference restatement.  On the GPU box the same tests run the
 * addon against the real libhgwarp.so instead.  Linked only into tests/napi_mock's library — never into the product,
 * which has no CPU path.  The host-only entry points (hg_delaunay, hg_png_*) are NOT doubled: they come from libhgwarp.so. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/hgwarp.h"

void orc_affine_from_triangles(const double *s, const double *d, float *out);
void orc_inverse_affine(const float *m, float *out);
void orc_projective_from_squares(const double *s, const double *d, double *out);
void orc_transform_limits(int kind, const void *matrix, double width, double height, double *out);
void orc_build_index_map(const float *pts, const uint32_t *tris, int32_t n_tris, double matrix_width, double y_offset,
                         int16_t *map, int64_t len);
void orc_piecewise_matrices(const float *src_pts, const float *dst_pts, const uint32_t *tris, int32_t n_tris, float *out);
void orc_warp_inverse_geometric(int kind, const uint8_t *img, int32_t W, int32_t H, const void *inv, int32_t xOff, int32_t yOff,
                                int32_t oW, int32_t oH, uint8_t *out, int threads);
void orc_warp_forward_geometric(int kind, const uint8_t *img, int32_t W, int32_t H, const void *fwd, int32_t xOff, int32_t yOff,
                                int32_t oW, int32_t oH, uint8_t *out);
void orc_warp_inverse_piecewise(const uint8_t *img, int32_t W, int32_t H, const int16_t *map, int64_t map_len, const float *inv,
                                int32_t n_tris, int32_t xOff, int32_t yOff, int32_t oW, int32_t oH, int32_t minSrcX,
                                int32_t minSrcY, uint8_t *out, int threads);
void orc_warp_forward_piecewise(const uint8_t *img, int32_t W, int32_t H, const int16_t *map, int64_t map_len, const float *fwd,
                                int32_t n_tris, int32_t xOff, int32_t yOff, int32_t oW, int32_t oH, int32_t minSrcX,
                                int32_t minSrcY, int32_t maxSrcX, int32_t maxSrcY, uint8_t *out);

struct hg_ctx {
    uint8_t *img;
    int W, H;
    float *src_pts;
    uint32_t *tris;
    int n_pts, n_tris;
    int16_t *last_map;   /* the map left by the last inverse piecewise warp (the aliasing forward read, Q8) */
    int64_t last_map_len;
    char err[256];
};

static int fail(hg_ctx *c, int code, const char *msg)
{
    if (c) snprintf(c->err, sizeof c->err, "%s", msg);
    return code;
}

int hg_ctx_create(int device, hg_ctx **out)
{
    if (!out || device != 0) return HG_ERR_INVALID;
    *out = (hg_ctx *)calloc(1, sizeof(hg_ctx));
    return *out ? HG_OK : HG_ERR_NOMEM;
}
int hg_ctx_destroy(hg_ctx *c)
{
    if (!c) return HG_ERR_INVALID;
    free(c->img); free(c->src_pts); free(c->tris); free(c->last_map); free(c);
    return HG_OK;
}
const char *hg_last_error(hg_ctx *c) { return c ? c->err : "create failed"; }

int hg_image_set(hg_ctx *c, const uint8_t *rgba, int w, int h)
{
    if (!c || !rgba || w < 1 || h < 1) return fail(c, HG_ERR_INVALID, "bad image");
    free(c->img);
    c->img = (uint8_t *)malloc((size_t)w * h * 4);
    memcpy(c->img, rgba, (size_t)w * h * 4);
    c->W = w; c->H = h;
    return HG_OK;
}

static void solve(int kind, const double *from, const double *to, void *m)
{
    if (kind == HG_AFFINE) orc_affine_from_triangles(from, to, (float *)m);
    else orc_projective_from_squares(from, to, (double *)m);
}

int hg_solve_with_limits(hg_ctx *c, int kind, const double *src, const double *dst, double w, double h, void *matrix_out,
                         double limits_out[4])
{
    if (!c || !src || !dst || !matrix_out || !limits_out) return fail(c, HG_ERR_INVALID, "NULL argument");
    solve(kind, src, dst, matrix_out);
    orc_transform_limits(kind, matrix_out, w, h, limits_out);
    return HG_OK;
}

static int need_image_and_window(hg_ctx *c, int o_w, int o_h, uint8_t *out_host)
{
    if (!c->img) return fail(c, HG_ERR_STATE, "no image set (hg_image_set)");
    if (o_w < 1 || o_h < 1) return fail(c, HG_ERR_INVALID, "output size must be >= 1x1");
    if (!out_host) return fail(c, HG_ERR_INVALID, "the CPU double only writes host buffers");
    return HG_OK;
}

int hg_warp_inverse_points(hg_ctx *c, int kind, const double *dst_pts, const double *src_pts, int x_off, int y_off, int o_w,
                           int o_h, uint8_t *out_host, void *out_dev)
{
    (void)out_dev;
    int st = need_image_and_window(c, o_w, o_h, out_host);
    if (st) return st;
    double m[8];
    solve(kind, dst_pts, src_pts, m);   /* calculateTransformMatrix(kind, dst, src), H.js:994 */
    orc_warp_inverse_geometric(kind, c->img, c->W, c->H, m, x_off, y_off, o_w, o_h, out_host, 1);
    return HG_OK;
}

int hg_warp_forward_matrix(hg_ctx *c, int kind, const void *fwd, int x_off, int y_off, int o_w, int o_h, uint8_t *out_host,
                           void *out_dev)
{
    (void)out_dev;
    int st = need_image_and_window(c, o_w, o_h, out_host);
    if (st) return st;
    orc_warp_forward_geometric(kind, c->img, c->W, c->H, fwd, x_off, y_off, o_w, o_h, out_host);
    return HG_OK;
}

int hg_piecewise_set_mesh(hg_ctx *c, const float *src_pts, int n_pts, const uint32_t *tris, int n_tris)
{
    if (!c || !src_pts || !tris || n_pts < 3 || n_tris < 0) return fail(c, HG_ERR_INVALID, "bad mesh");
    free(c->src_pts); free(c->tris);
    c->src_pts = (float *)malloc(sizeof(float) * 2 * (size_t)n_pts);
    c->tris = (uint32_t *)malloc(sizeof(uint32_t) * 3 * (size_t)(n_tris ? n_tris : 1));
    memcpy(c->src_pts, src_pts, sizeof(float) * 2 * (size_t)n_pts);
    memcpy(c->tris, tris, sizeof(uint32_t) * 3 * (size_t)n_tris);
    c->n_pts = n_pts; c->n_tris = n_tris;
    return HG_OK;
}

int hg_piecewise_mesh_size(hg_ctx *c, int *n_pts, int *n_tris)
{
    if (!c || !n_pts || !n_tris) return HG_ERR_INVALID;
    *n_pts = c->n_pts;
    *n_tris = c->n_tris;
    return HG_OK;
}

int hg_piecewise_matrices(hg_ctx *c, const float *dst_pts, float *fwd_out, float *inv_out)
{
    if (!c || !c->src_pts || !dst_pts) return fail(c, HG_ERR_STATE, "no mesh");
    float *fwd = (float *)malloc(sizeof(float) * 6 * (size_t)(c->n_tris ? c->n_tris : 1));
    orc_piecewise_matrices(c->src_pts, dst_pts, c->tris, c->n_tris, fwd);
    if (fwd_out) memcpy(fwd_out, fwd, sizeof(float) * 6 * (size_t)c->n_tris);
    if (inv_out)
        for (int t = 0; t < c->n_tris; ++t) orc_inverse_affine(fwd + 6 * t, inv_out + 6 * t);
    free(fwd);
    return HG_OK;
}

int hg_warp_piecewise_inverse(hg_ctx *c, const float *dst_pts, int x_off, int y_off, int o_w, int o_h, int min_src_x,
                              int min_src_y, uint8_t *out_host, void *out_dev)
{
    (void)out_dev;
    int st = need_image_and_window(c, o_w, o_h, out_host);
    if (st) return st;
    if (!c->src_pts) return fail(c, HG_ERR_STATE, "no mesh");
    const size_t nt = (size_t)(c->n_tris ? c->n_tris : 1);
    float *fwd = (float *)malloc(sizeof(float) * 6 * nt), *inv = (float *)malloc(sizeof(float) * 6 * nt);
    orc_piecewise_matrices(c->src_pts, dst_pts, c->tris, c->n_tris, fwd);
    for (int t = 0; t < c->n_tris; ++t) orc_inverse_affine(fwd + 6 * t, inv + 6 * t);
    free(c->last_map);
    c->last_map_len = (int64_t)o_w * o_h;
    c->last_map = (int16_t *)malloc(sizeof(int16_t) * (size_t)c->last_map_len);
    orc_build_index_map(dst_pts, c->tris, c->n_tris, (double)o_w, (double)y_off, c->last_map, c->last_map_len);
    orc_warp_inverse_piecewise(c->img, c->W, c->H, c->last_map, c->last_map_len, inv, c->n_tris, x_off, y_off, o_w, o_h,
                               min_src_x, min_src_y, out_host, 1);
    free(fwd); free(inv);
    return HG_OK;
}

int hg_warp_piecewise_forward(hg_ctx *c, const float *dst_pts, int x_off, int y_off, int o_w, int o_h, int min_src_x,
                              int min_src_y, int max_src_x, int max_src_y, int use_inverse_map, uint8_t *out_host, void *out_dev)
{
    (void)out_dev;
    int st = need_image_and_window(c, o_w, o_h, out_host);
    if (st) return st;
    if (!c->src_pts) return fail(c, HG_ERR_STATE, "no mesh");
    const size_t nt = (size_t)(c->n_tris ? c->n_tris : 1);
    float *fwd = (float *)malloc(sizeof(float) * 6 * nt);
    orc_piecewise_matrices(c->src_pts, dst_pts, c->tris, c->n_tris, fwd);
    int16_t *map = c->last_map;
    int64_t len = c->last_map_len;
    int16_t *own = NULL;
    if (!use_inverse_map) {
        const int64_t mw = (int64_t)max_src_x - min_src_x;
        len = mw * ((int64_t)max_src_y - min_src_y);
        own = map = (int16_t *)malloc(sizeof(int16_t) * (size_t)(len > 0 ? len : 1));
        orc_build_index_map(c->src_pts, c->tris, c->n_tris, (double)mw, (double)min_src_y, map, len > 0 ? len : 0);
    }
    orc_warp_forward_piecewise(c->img, c->W, c->H, map, len, fwd, c->n_tris, x_off, y_off, o_w, o_h, min_src_x, min_src_y,
                               max_src_x, max_src_y, out_host);
    free(own); free(fwd);
    return HG_OK;
}
